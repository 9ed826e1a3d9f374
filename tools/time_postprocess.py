"""Device time of remove_all_but_the_largest_connected_component on a 512 x 512 x 256 label volume (13 classes one by one,
and all foreground classes as one region)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from hybrid_ctunet_b200.postprocess import remove_all_but_the_largest_connected_component as ours
g = torch.Generator(device="cuda").manual_seed(0)
img = torch.zeros(512, 512, 256, dtype=torch.int64, device="cuda")
for c in range(1, 14):
    s = torch.nn.functional.avg_pool3d(torch.rand(1, 1, 512, 512, 256, device="cuda", generator=g), 3, stride=1, padding=1)[0, 0]
    img[(s > 0.56) & (img == 0)] = c
for name, classes in (("13 classes", list(range(1, 14))), ("foreground as one region", [tuple(range(1, 14))])):
    ours(img, classes, 1.0); torch.cuda.synchronize()
    t0 = time.time(); out, rem, kept = ours(img, classes, 1.0); torch.cuda.synchronize(); dt = time.time() - t0
    print(f"GPU {name}: {dt * 1e3:.1f} ms  (removed {int((out != img).sum())} voxels)", flush=True)
# (The reference's host loop — one full-volume pass per OBJECT, test_CTUNet_final.py:164-166 — did not finish one class of a
# 256 x 256 x 128 sub-volume of this noise-like test volume within 10 minutes on the GPU box's host: not timed here.)
