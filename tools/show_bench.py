"""Pretty-print one bench.py JSON line: headline numbers and the per-kernel-class roofline table."""
import json
import sys

d = json.load(open(sys.argv[1]))
print(f"{d['ms_per_step']:.2f} ms/step  {d['value']:.2f} {d['unit']}  e2e {d['e2e']['value']:.2f}  n_gpus {d['n_gpus']}")
for k in ("sliding_window", "hybrid", "config4"):
    if d.get(k):
        v = d[k]
        print(f"{k}: {v['value']:.4f} {v['unit']}  ({v.get('ms_per_volume', v.get('ms_per_step')):.1f} ms)" +
              (f"  e2e {v['e2e']['value']:.4f}" if "e2e" in v else ""))
if d.get("library_gpu"):
    print("library_gpu:", json.dumps(d["library_gpu"])[:600])
r = d.get("roofline")
if r:
    print({k: v for k, v in r.items() if k not in ("classes", "how", "kernel")})
    for c in r["classes"]:
        print(f"{c['ms_per_step']:7.2f} ms {c['share_of_step'] * 100:5.1f}%  {c['achieved']:8.1f} {c['unit']:8s} frac {c['frac']:.3f} "
              f"x{c['launches_per_step']:4d}  {c['class'][:72]}")
        for t in c.get("top_shapes", []):
            print("        ", t)
