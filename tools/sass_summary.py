"""profiles/rNN_sass_summary.txt: mnemonic counts per kernel of the in-tree library + the consumer loop of the default
merge kernel (what proves TMA bulk copies / mbarriers / 256-bit loads are in the binary).  Runs without a GPU.

    python tools/sass_summary.py [profiles/r02_sass_summary.txt]
"""
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "camera_linearity_b200" / "libcamlin_b200.so"
SHOW = ["BAR.SYNC", "MUFU", "LDG", "LDGSTS", "STG", "ATOM", "RED", "DADD", "DFMA", "DMUL", "LDS.64", "LDS.128", "SYNCS.PHASECHK",
        "SYNCS.ARRIVE", "UBLKCP", "MEMBAR", "ENL2.256"]
EXCERPT_KERNEL = "merge_stream_kernelILb0"


def main():
    out = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "profiles" / "r02_sass_summary.txt"
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels, name = {}, None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = m.group(1)
            kernels[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            kernels[name].append(ln)
    lines = ["# cuobjdump -sass camera_linearity_b200/libcamlin_b200.so (sm_100a, the in-tree build of this commit): "
             "mnemonic counts per kernel (tools/sass_summary.py)",
             "# UBLKCP = cp.async.bulk (TMA 1-D bulk copy), SYNCS.* = mbarrier operations, LDGSTS = cp.async, "
             "ENL2.256 = 256-bit global load", ""]
    for k in sorted(kernels):
        cnt = Counter()
        for ln in kernels[k]:
            body = re.sub(r"^\s*/\*[0-9a-f]+\*/\s+(@!?U?P\d\s+)?", "", ln)
            op = body.split()[0].rstrip(";")
            for key in SHOW:
                if key == "ENL2.256":
                    if "ENL2.256" in op:
                        cnt[key] += 1
                elif op == key or op.startswith(key + "."):
                    cnt[key] += 1
        short = re.sub(r"^_ZN2cl\d*", "", k)
        short = re.sub(r"^_GLOBAL__N__[0-9a-f]+_\d+_", "", short)
        lines.append(f"{short[:110]:112s} {len(kernels[k]):5d} instr  " + " ".join(f"{a}={b}" for a, b in cnt.items()))
    ex = next((k for k in kernels if EXCERPT_KERNEL in k), None)
    if ex:
        body = kernels[ex]
        # the consumer loop: from the first LDS.128 (table row) to the stage release that follows it
        first = next((i for i, ln in enumerate(body) if "LDS.128" in ln), 0)
        last = next((i for i in range(first, len(body)) if "SYNCS.ARRIVE" in body[i]), min(first + 150, len(body) - 1))
        lines += ["", f"# consumer loop of {EXCERPT_KERNEL} (one exposure of one pixel: 3 channels), SASS lines "
                      f"{max(first - 30, 0)}..{last + 12}:"]
        lines += body[max(first - 30, 0):last + 12]
    out.write_text("\n".join(lines) + "\n")
    print(f"{out}: {len(kernels)} kernels")


if __name__ == "__main__":
    main()
