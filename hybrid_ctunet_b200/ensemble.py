"""Mask-complementation ensemble of the reference's evaluation scripts on the device.

test_CTUNet.py:236-251 / test_CTUNet_final.py:547-552 / trainer_CTUNet.py:287-299: softmax of each blended head,
their mean, three argmax masks, then `dice(pred == i, label == i)` for the 13 organ classes (utils/utils.py:16-22:
2*|x*y| / (|x| + |y|), 0 when the label class is empty).  `ensemble_masks` does all of it in one kernel
(ctu_ensemble_argmax) over the two fp32 logit volumes; only the uint8 masks and 126 counters leave the GPU.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import lib as _lib


def ensemble_masks(pred1: torch.Tensor, pred2: torch.Tensor, labels: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """pred1 / pred2: fp32 [C, X, Y, Z] (or [1, C, X, Y, Z]) blended logits; labels: [X, Y, Z] (any leading 1s).
    Returns uint8 masks "ensemble", "head1", "head2" and, with labels, float64 "dice" [3, C] (rows as the masks)."""
    lib = _lib.require_device()
    p1 = pred1.reshape(pred1.shape[-4:]).float().contiguous()
    p2 = pred2.reshape(pred2.shape[-4:]).float().contiguous()
    if p1.shape != p2.shape or not p1.is_cuda:
        raise ValueError("pred1 / pred2 must be CUDA tensors of the same [C, X, Y, Z] shape")
    C = p1.shape[0]
    V = p1[0].numel()
    out = {k: torch.empty(p1.shape[1:], dtype=torch.uint8, device=p1.device) for k in ("ensemble", "head1", "head2")}
    lab = counts = None
    if labels is not None:
        lab = labels.reshape(p1.shape[1:]).float().contiguous()
        counts = torch.zeros(3, C, 3, dtype=torch.int64, device=p1.device)
    stream = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.ctu_ensemble_argmax(p1.data_ptr(), p2.data_ptr(), C, V, out["ensemble"].data_ptr(), out["head1"].data_ptr(),
                                       out["head2"].data_ptr(), None if lab is None else lab.data_ptr(),
                                       None if counts is None else counts.data_ptr(), stream), "ctu_ensemble_argmax")
    if counts is not None:
        c = counts.double()
        dice = torch.where(c[..., 2] > 0, 2.0 * c[..., 0] / (c[..., 1] + c[..., 2]).clamp_min(1.0), torch.zeros_like(c[..., 0]))
        out["dice"], out["counts"] = dice, counts
    return out


def hybrid_ctunet_inference(inputs: torch.Tensor, ctunet, tunet, roi_size=(96, 96, 96), sw_batch_size: int = 4,
                            labels: Optional[torch.Tensor] = None, shard_group=None, invert=None) -> Dict[str, torch.Tensor]:
    """The "Hybrid-CTUNet" configuration of test_CTUNet_final.py:539-552 for one volume [1, 1, X, Y, Z]: CTUNet's
    ResNet-branch head blended at overlap 0.5, an independently trained TUNet's first head blended at overlap 0.7
    (one-head sliding window), then the mask-complementation ensemble.  Windows are sharded over `shard_group`.
    `invert` (an `invert.InvertGeometry`): the scripts' `Invertd` step between the two (test_CTUNet_final.py:541-545) — the
    masks (and `labels`) are then on the voxel grid of the file, computed by the fused inverse + ensemble kernel."""
    from .sliding_window import sliding_window_inference_one_head
    with torch.no_grad():
        # only CTUNet's ResNet-branch head is used (`[0]` of the two blended heads in the reference script): blend just
        # that one — half the accumulator traffic and, when sharded, half the bytes on the wire
        p1 = sliding_window_inference_one_head(inputs, roi_size, sw_batch_size, lambda w: (ctunet(w)[0][0],), overlap=0.5,
                                               mode="gaussian", shard_group=shard_group)
        p2 = sliding_window_inference_one_head(inputs, roi_size, sw_batch_size, tunet, overlap=0.7, mode="gaussian",
                                               shard_group=shard_group)
        if invert is not None:
            from .invert import invert_ensemble_masks
            return invert_ensemble_masks(p1[0], p2[0], invert, labels)
        return ensemble_masks(p1[0], p2[0], labels)


def evaluate_cases(cases, ctunet, tunet, roi_size=(96, 96, 96), sw_batch_size: int = 4, postprocess: bool = True,
                   dice_threshold: float = 0.0, advanced_postprocessing: bool = True, shard_group=None) -> Dict[str, object]:
    """The evaluation loop of test_CTUNet_final.py:527-655 with every per-voxel step on the device.  `cases`: an iterable of
    dicts with "image" [1, 1, x, y, z] (the loader's resampled, cropped tensor), "label" (the file's own label volume
    [X0, Y0, Z0], any leading 1s), "geometry" (an `invert.InvertGeometry`, or None when the image already is on the label's
    grid) and "volume_per_voxel" (float, for the size thresholds).  Per case: both sliding windows, `Invertd` + softmax
    average + three argmax masks + 13-class Dice in one kernel; then, over all cases, `determine_postprocessing`
    (largest-connected-component rule) as the script's closing step.  Returns the per-case masks (uint8, CUDA), the three
    Dice tables [case][13] (ResNet head / TUNet head / ensemble, the script's `dice_list_case_res / _vit / dice_list_case`),
    their organ means and — with `postprocess` — the post-processed masks and their mean Dice (`com_dice`)."""
    import numpy as np
    from .postprocess import com_dice, determine_postprocessing
    masks, labels, vpv = [], [], []
    tables = {"res": [], "vit": [], "ensemble": []}
    for case in cases:
        lab = case["label"]
        lab = lab.reshape(lab.shape[-3:])
        out = hybrid_ctunet_inference(case["image"], ctunet, tunet, roi_size, sw_batch_size, labels=lab,
                                      shard_group=shard_group, invert=case.get("geometry"))
        dice = out["dice"][:, 1:].cpu().numpy()           # rows: ensemble, head 1 (ResNet branch), head 2 (TUNet); organs 1..13
        tables["ensemble"].append(dice[0])
        tables["res"].append(dice[1])
        tables["vit"].append(dice[2])
        masks.append(out["ensemble"])
        labels.append(lab)
        vpv.append(float(case.get("volume_per_voxel", 1.0)))
    res: Dict[str, object] = {"masks": masks, "dice": {k: np.asarray(v) for k, v in tables.items()},
                              "mean_organ_dice": {k: np.mean(np.asarray(v), axis=0) for k, v in tables.items()}}
    if postprocess and masks:
        post = determine_postprocessing(masks, labels, vpv, dice_threshold=dice_threshold,
                                        advanced_postprocessing=advanced_postprocessing)
        res["masks_postprocessed"] = post
        res["mean_organ_dice_postprocessed"] = np.asarray(com_dice(post, labels))
    return res
