"""Per-kernel totals of ONE training step out of an ncu launch list of bench.py (`--metrics gpu__time_duration.sum --csv`):
from the last `pack_weights_kernel` launch before the last `unpack_grads_kernel` up to that launch (forward + loss + backward).  Prints class and per-kernel totals."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr, data = rows[hi], rows[hi + 1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = [(re.sub(r"\(.*", "", r[ki]).replace("ctu::", "").replace("void ", ""), float(r[vi].replace(",", "")) / 1e3) for r in data]
unp = [i for i, (n, _) in enumerate(seq) if n.startswith("unpack_grads_kernel")]
hi = unp[-1] + 1
lo = max(i for i, (n, _) in enumerate(seq) if n.startswith("pack_weights_kernel") and i < hi)
step = seq[lo:hi]
tot = sum(t for _, t in step)
print(f"one training step (launches {lo}..{hi}): {len(step)} launches, {tot / 1e3:.2f} ms of kernels serialised")
CLASSES = [("tcgen05 conv/gemm/dgrad", ("umma_gemm_kernel", "conv3_halo_kernel", "conv3_halo_x2_kernel", "ffn_fused_kernel")), ("tcgen05 wgrad", ("umma_wgrad_kernel", "wgrad_halo_kernel")),
           ("InstanceNorm", ("in_", "stats_fold")), ("attention", ("attention",)), ("LayerNorm", ("layernorm", "patchify")),
           ("pack/unpack/AdamW", ("pack_weights", "unpack_grads", "adamw")), ("gelu/pwa", ("gelu", "pwa_")),
           ("loss", ("dice_ce", "gather3d")),
           ("colsum/accumulate/cast/layout", ("colsum", "accumulate", "cast_", "space_to_depth", "cf_to_cl", "subsample", "im2col", "conv_cin1")),
           ("head backward (fused CUDA-core)", ("head_bwd",))]
cl = collections.OrderedDict((c, 0.0) for c, _ in CLASSES)
cl["torch (loss glue, fills, copies)"] = 0.0
agg = collections.OrderedDict()
for n, t in step:
    for c, pre in CLASSES:
        if any(n.startswith(p) for p in pre):
            cl[c] += t
            break
    else:
        cl["torch (loss glue, fills, copies)"] += t
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += t
for c, t in cl.items():
    print(f"{c:36s} {t / 1e3:7.2f} ms {100 * t / tot:5.1f}%")
print()
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{t:9.1f} us {100 * t / tot:5.1f}%  x{n:4d}  {k}")
