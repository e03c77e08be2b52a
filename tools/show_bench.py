"""Print the interesting parts of a bench.py output file (the JSON line may be preceded by library banners)."""
import json
import sys

for path in sys.argv[1:]:
    line = [l for l in open(path).read().splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    print(f"== {path}: n_gpus {d['n_gpus']}  value {d['value']:.1f} {d['unit']}  ms/step {d['ms_per_step']:.4f}  "
          f"frac {d['roofline']['frac']:.4f}")
    e = d["e2e"]
    print(f"   e2e {e['value']:.2f} ({e['ms_per_step']:.1f} ms/step)", end="")
    sv = d.get("std_table_variant")
    if sv and "e2e" in sv:
        print(f" | STD-table: {sv['value']:.1f} dev ({sv['ms_per_step']:.4f} ms), e2e {sv['e2e']['value']:.2f} "
              f"({sv['e2e']['ms_per_step']:.2f} ms/step)", end="")
    print()
    k4 = d.get("k4_icrf_fit") or {}
    for name in ("nostd", "std"):
        if name in k4:
            r = k4[name]
            g = r.get("de_generation", {})
            print(f"   k4 {name}: {r['ms_per_population']:.4f} ms/pop = {r['evals/s']:.0f} evals/s ({r['exchange']}); "
                  f"DE generation {g.get('ms', float('nan')):.4f} ms = {g.get('evals/s', float('nan')):.0f} evals/s"
                  + (f"; cpu {r['cpu_baseline']['value']:.2f} evals/s" if 'cpu_baseline' in r else "")
                  + (f"; NCCL variant {r['nccl_allreduce_variant']['ms_per_population']:.4f} ms/pop, DE "
                     f"{r['nccl_allreduce_variant']['de_generation_ms']:.4f} ms" if 'nccl_allreduce_variant' in r else ""))
    if "error" in k4:
        print("   k4 error:", k4["error"])
    c5 = d.get("cfg5_sharded_batch") or {}
    for name in ("f64_std", "std_table"):
        if name in c5:
            r = c5[name]
            print(f"   cfg5 {name}: {r['value']:.1f} Gpix*exp/s, {r['ms_per_stack']:.3f} ms/stack, frac {r['frac_of_hbm_peak']:.3f}, "
                  f"{r['stacks_per_rank_per_step']} stacks/rank/step")
    if "error" in c5:
        print("   cfg5 error:", c5["error"])
    for k, v in (d.get("extra") or {}).items():
        if isinstance(v, dict):
            print("   ", k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a not in ("shape", "cpu_baseline")})
        else:
            print("   ", k, v)
    if d.get("cpu_baseline"):
        print("   cpu:", d["cpu_baseline"]["value"])
    print("   clocks:", d.get("clocks"))
