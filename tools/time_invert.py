"""Times the device `Invertd` (+ fused ensemble) on a 512 x 512 x 147 scan (0.76 x 0.76 x 3.0 mm, resampled by the loader to
1.5 x 1.5 x 2.0 mm): CUDA events around each call, after warm-up.  Usage: python tools/time_invert.py [--host] (--host also
times torch's float64 affine_grid + grid_sample on the host cores for ONE class plane: the reference's inner operation)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hybrid_ctunet_b200.ensemble import ensemble_masks  # noqa: E402
from hybrid_ctunet_b200.invert import InvertGeometry, invert_ensemble_masks, invert_pred  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return min(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))


def main():
    shape = (512, 512, 147)
    aff = np.diag([-0.76, 0.76, 3.0, 1.0])
    aff[:3, 3] = (190.0, -170.0, -300.0)
    ps = InvertGeometry.from_file(aff, shape, (1.5, 1.5, 2.0), (0, 0, 0), (1, 1, 1)).pad_size
    g = InvertGeometry.from_file(aff, shape, (1.5, 1.5, 2.0), (11, 17, 6), (ps[0] - 9, ps[1] - 20, ps[2] - 4))
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(0)
    p1 = torch.randn((14,) + g.pred_size, generator=gen, device=dev)
    p2 = torch.randn((14,) + g.pred_size, generator=gen, device=dev)
    V, P = int(np.prod(shape)), int(np.prod(g.pred_size))
    res = {"out_size": shape, "pred_size": g.pred_size, "pad_size": g.pad_size}
    t = timed(lambda: invert_pred(p1, g))
    res["invert_pred_ms"] = round(t, 3)
    res["invert_pred_GBs"] = round((14 * P * 4 + 14 * V * 4) / t / 1e6, 1)
    t = timed(lambda: invert_ensemble_masks(p1, p2, g))
    res["fused_invert_ensemble_ms"] = round(t, 3)
    res["fused_GBs"] = round((2 * 14 * P * 4 + 3 * V) / t / 1e6, 1)
    i1, i2 = invert_pred(p1, g), invert_pred(p2, g)
    t2 = timed(lambda: ensemble_masks(i1, i2))
    res["two_step_ms"] = round(2 * res["invert_pred_ms"] + t2, 3)
    if "--host" in sys.argv:
        # the reference's inner operation for ONE of the 14 class planes: float64 affine_grid + grid_sample on the host cores
        plane = torch.zeros((1, 1) + tuple(g.pad_size), dtype=torch.float64)
        theta = torch.tensor([[[0.98, 0, 0, 0], [0, 0.98, 0, 0], [0, 0, 0.99, 0]]], dtype=torch.float64)
        t0 = time.perf_counter()
        grid = torch.nn.functional.affine_grid(theta, [1, 1] + list(shape), align_corners=False)
        torch.nn.functional.grid_sample(plane, grid, mode="bilinear", padding_mode="border", align_corners=False)
        res["host_grid_sample_one_plane_s"] = round(time.perf_counter() - t0, 2)
        res["host_threads"] = torch.get_num_threads()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
