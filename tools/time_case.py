"""Per-case latency of the evaluation script's loop body (test_CTUNet_final.py:527-566) on one GPU, everything on the device:
a 512 x 512 x 147 scan at 0.76 x 0.76 x 3.0 mm that the loader resampled to 1.5 x 1.5 x 2.0 mm and cropped to 240 x 223 x 210
-> CTUNet head 0 blended at overlap 0.5 + TUNet blended at overlap 0.7 -> Invertd of both + softmax mean + argmax masks + Dice
(one kernel) -> largest-connected-component filter of the 13 organ classes.  Synthetic image / labels, random-init weights.
The host->device copy of the image and the device->host copy of the final uint8 mask are inside the timed region."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hybrid_ctunet_b200.ensemble import hybrid_ctunet_inference  # noqa: E402
from hybrid_ctunet_b200.invert import InvertGeometry  # noqa: E402
from hybrid_ctunet_b200.networks.hybrid_CTUNet import CTUNet, TUNet  # noqa: E402
from hybrid_ctunet_b200.postprocess import remove_all_but_the_largest_connected_component  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
kw = dict(in_channels=1, dim_conv_stem=64, out_channels=14, img_size=(96, 96), frames=96, patch_frame=8)
torch.manual_seed(0)
ctunet = CTUNet(model_depth=101, **kw).to(dev).eval().enable_cuda_graph()
torch.manual_seed(5)
tunet = TUNet(**kw).to(dev).eval().enable_cuda_graph()

shape = (512, 512, 147)
aff = np.diag([-0.76, 0.76, 3.0, 1.0])
aff[:3, 3] = (190.0, -170.0, -300.0)
ps = InvertGeometry.from_file(aff, shape, (1.5, 1.5, 2.0), (0, 0, 0), (1, 1, 1)).pad_size
geom = InvertGeometry.from_file(aff, shape, (1.5, 1.5, 2.0), (11, 17, 6), (ps[0] - 9, ps[1] - 20, ps[2] - 4))
torch.manual_seed(2)
image = torch.rand((1, 1) + geom.pred_size).pin_memory()
label = torch.randint(0, 14, shape, dtype=torch.uint8).pin_memory()
vpv = 0.76 * 0.76 * 3.0


def case():
    img = image.to(dev, non_blocking=True)
    lab = label.to(dev, non_blocking=True)
    stages = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    stages[0].record()
    out = hybrid_ctunet_inference(img, ctunet, tunet, labels=lab, invert=geom)
    stages[1].record()
    post, _, _ = remove_all_but_the_largest_connected_component(out["ensemble"], list(range(1, 14)), vpv)
    stages[2].record()
    mask = post.cpu()                       # what nib.save would write
    dice = out["dice"].cpu()
    return mask, dice, stages


case()                                      # warm-up: graph captures, allocator
torch.cuda.synchronize()
t0 = time.perf_counter()
mask, dice, st = case()
torch.cuda.synchronize()
wall = time.perf_counter() - t0
windows = {"ctunet@0.5": 4 * 4 * 4, "tunet@0.7": 7 * 6 * 6}
print(json.dumps({"case_wall_s": round(wall, 4), "inference_invert_ensemble_ms": round(st[0].elapsed_time(st[1]), 2),
                  "connected_components_13_classes_ms": round(st[1].elapsed_time(st[2]), 2), "windows": windows,
                  "image": list(geom.pred_size), "file_grid": list(shape), "mask_shape": list(mask.shape),
                  "mean_dice_ensemble": float(dice[0, 1:].mean())}))
