"""Summarise an .ncu-rep (one kernel) into a small text file for profiles/.

    python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_xxx.txt ["note"] [extra ncu import args ...]

e.g. ``-k regex:dark_scan -c 1`` to pick one kernel out of a multi-kernel report.
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__cycles_elapsed.avg.per_second",
]
STALLS = "smsp__average_warps_issue_stalled_"


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    extra = sys.argv[4:]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", *extra], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = [f"# {rep}", f"# {note}", ""]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        lines.append(f"kernel: {d.get('Kernel Name')}")
        for k in KEYS:
            if k in d:
                lines.append(f"  {k:85s} {d[k]:>18s} {u.get(k, '')}")
        for k in hdr:
            if k.startswith(STALLS) and k.endswith("per_issue_active.ratio") and float(d[k] or 0) >= 0.05:
                lines.append(f"  {k:85s} {d[k]:>18s}")
        lines.append("")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", *extra], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    if len(rows) > 2:
        h = rows[1]
        ix = {name: i for i, name in enumerate(h)}
        data = [r for r in rows[2:] if len(r) == len(h) and r[ix["# Samples"]].isdigit()]   # (headers repeat per kernel)
        tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
        lines.append(f"top stall sites (of {tot} samples, {len(data)} SASS instructions):")
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:15]:
            lines.append(f"  {100 * int(r[ix['# Samples']]) / tot:5.1f}%  long_sb={r[ix['stall_long_sb']]:>6s} "
                         f"short_sb={r[ix['stall_short_sb']]:>6s} wait={r[ix['stall_wait']]:>6s}  {r[ix['Source']].strip()[:80]}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
