"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and the top launches."""
import collections, csv, re, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
agg, seq, tot = collections.OrderedDict(), [], 0.0
for r in data:
    name = re.sub(r"\(.*", "", r[ki]).replace("ctu::", "").replace("void ", "")
    t = float(r[vi].replace(",", "")) / 1e3
    seq.append((name, r[gi], t))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t; tot += t
print(f"total {tot:.1f} us over {len(seq)} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.1f} us {100 * t / tot:5.1f}%  x{n:4d}  {k}")
print()
for name, g, t in sorted(seq, key=lambda s: -s[2])[:top]:
    print(f"{t:9.1f} us  grid {g:>18s}  {name}")
if len(sys.argv) > 3:  # dump the sequence
    for i, (name, g, t) in enumerate(seq):
        print(i, f"{t:9.1f}", g, name)
