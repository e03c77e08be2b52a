"""profiles/r02_traffic.json from the ncu launch lists (what bench.py's `roofline.traffic` reads).

    python tools/traffic_from_launches.py profiles/r02_launches_cfg2.csv [profiles/r02_launches_cfg2_lut.csv]

Per kernel of our library (namespace cl): launches, average duration, average DRAM bytes read / written; the
merge kernel is the one with the largest total time, the others are listed beside it with their share of the step.
"""
import csv
import json
import re
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OURS = re.compile(r"cl::(?:<unnamed>::)?(\w+)")


def per_kernel(path):
    rows = defaultdict(lambda: defaultdict(dict))          # name -> launch id -> metric -> value
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        m = OURS.search(r["Kernel Name"])
        if not m:
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        else:
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        rows[m.group(1)][r["ID"]][r["Metric Name"]] = v
    out = {}
    for name, launches in rows.items():
        n = len(launches)
        out[name] = {"launches": n,
                     "avg_us": sum(x.get("gpu__time_duration.sum", 0.0) for x in launches.values()) / n,
                     "dram_bytes_read": sum(x.get("dram__bytes_read.sum", 0.0) for x in launches.values()) / n,
                     "dram_bytes_write": sum(x.get("dram__bytes_write.sum", 0.0) for x in launches.values()) / n}
    return out


def entry(kernels, merge_prefix, step_kernels):
    merge = max((k for k in kernels if k.startswith(merge_prefix)), key=lambda k: kernels[k]["avg_us"] * kernels[k]["launches"])
    n_steps = kernels[merge]["launches"]
    step = {k: v for k, v in kernels.items() if k == merge or k in step_kernels}
    # every helper is launched once per cl_hdr_merge / cl_flat_roi_means call (the capture also holds the calls of the
    # STD-table variant, so launch counts differ between kernels): per step = per launch
    per_step_us = {k: v["avg_us"] for k, v in step.items()}
    total = sum(per_step_us.values())
    return {"kernel": merge, "launches_in_capture": n_steps,
            "dram_bytes_read": kernels[merge]["dram_bytes_read"], "dram_bytes_write": kernels[merge]["dram_bytes_write"],
            "avg_us": {k: round(v, 2) for k, v in per_step_us.items()},
            "share_of_step": {k: round(v / total, 4) for k, v in per_step_us.items()},
            "other_kernels": {k: v["dram_bytes_read"] + v["dram_bytes_write"] for k, v in step.items() if k != merge}}


def main():
    src = sys.argv[1:] or [str(ROOT / "profiles" / "r02_launches_cfg2.csv")]
    helpers = {"roi_partial_kernel", "dark_scan_kernel", "merge_fixup_kernel", "merge_generic_kernel"}
    k = per_kernel(src[0])
    doc = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                     "python bench.py --steps 2 --warmup 3 --no-cpu --no-extra (" + ", ".join(Path(s).name for s in src) +
                     "; per launch, averaged over the launches of each kernel; tools/traffic_from_launches.py)",
           "cfg2": entry({n: v for n, v in k.items() if "_lut_" not in n}, "merge_s", helpers)}
    lut = {n: v for n, v in k.items() if n.startswith("merge_s") and "_lut_" in n}      # the STD-table variant's kernel
    if lut:
        name = max(lut, key=lambda n: lut[n]["avg_us"] * lut[n]["launches"])
        doc["cfg2_std_table"] = {"kernel": name, "dram_bytes_read": lut[name]["dram_bytes_read"],
                                 "dram_bytes_write": lut[name]["dram_bytes_write"], "avg_us": round(lut[name]["avg_us"], 2)}
    doc["all_kernels"] = {n: {kk: round(vv, 2) for kk, vv in v.items()} for n, v in sorted(k.items())}
    (ROOT / "profiles" / "r02_traffic.json").write_text(json.dumps(doc, indent=1) + "\n")
    print(json.dumps(doc["cfg2"], indent=1))


if __name__ == "__main__":
    main()
