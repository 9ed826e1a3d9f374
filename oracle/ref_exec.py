"""Run the reference's OWN trainer functions in this container (TEST INFRASTRUCTURE ONLY, build container only).

`/root/reference/trainer_CTUNet.py` and `trainer_CUNet.py` cannot be imported: their module headers pull in
tensorboardX and large parts of MONAI (trainer_CTUNet.py:19-39), neither installed nor installable here.  Their
function BODIES are plain torch/numpy, so this module parses the files with `ast`, takes the requested top-level
function / class definitions as they are — no source text is copied into the repo, nothing is edited — and executes
them in a namespace where the handful of MONAI names they reference are bound to the restatements in
oracle/sliding_window_oracle.py (MONAI 0.7.0 is an absent third-party dependency; its helpers stay restated, but the
loop, the 14-channel count maps, the blend and the crop that run are the reference's).

Used by tests/test_reference_exec.py (pins oracle/sliding_window_oracle.py and the drop-in's host logic against the
reference itself) and by tests/golden/make_golden.py (fixtures for the GPU box, where /root/reference does not exist).
"""
from __future__ import annotations

import ast
import enum
import os
from typing import Any, Callable, Dict, Iterable, List, Sequence, Tuple, Union

import numpy as np
import torch
import torch.nn.functional as F

from . import sliding_window_oracle as SO
from .ref_import import REFERENCE_ROOT


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "trainer_CTUNet.py"))


class BlendMode(enum.Enum):      # monai.utils.enums.BlendMode (0.7.0)
    CONSTANT = "constant"
    GAUSSIAN = "gaussian"


class PytorchPadMode(enum.Enum):  # monai.utils.enums.PytorchPadMode (0.7.0)
    CONSTANT = "constant"
    REFLECT = "reflect"
    REPLICATE = "replicate"
    CIRCULAR = "circular"


def look_up_option(opt_str, supported, default="no_default"):
    """monai.utils.look_up_option for Enum `supported`: the member whose value (or the member itself) matches."""
    if isinstance(opt_str, supported):
        return opt_str
    if isinstance(opt_str, str):
        opt_str = opt_str.strip()
    for m in supported:
        if m.value == opt_str:
            return m
    raise ValueError(f"Unsupported option '{opt_str}', available options are {[m.value for m in supported]}.")


def _compute_importance_map(patch_size, mode=BlendMode.CONSTANT, sigma_scale=0.125, device="cpu"):
    mode = look_up_option(mode, BlendMode)
    return SO.compute_importance_map(tuple(patch_size), mode=mode.value, sigma_scale=sigma_scale).to(device)


def monai_namespace() -> Dict[str, Any]:
    """Globals the extracted functions see: torch / numpy / typing plus the restated MONAI 0.7.0 helpers."""
    import scipy.ndimage as ndimage
    return dict(torch=torch, F=F, np=np, ndimage=ndimage, Any=Any, Callable=Callable, List=List, Sequence=Sequence,
                Tuple=Tuple, Union=Union, BlendMode=BlendMode, PytorchPadMode=PytorchPadMode,
                look_up_option=look_up_option, fall_back_tuple=SO.fall_back_tuple,
                get_valid_patch_size=SO.get_valid_patch_size, dense_patch_slices=SO.dense_patch_slices,
                compute_importance_map=_compute_importance_map)


def extract(rel_path: str, names: Iterable[str], extra_globals: Dict[str, Any] = None) -> Dict[str, Any]:
    """Compile the top-level definitions `names` of /root/reference/<rel_path> unchanged and return them."""
    path = os.path.join(REFERENCE_ROOT, rel_path)
    with open(path, "r") as fh:
        tree = ast.parse(fh.read(), filename=path)
    want = set(names)
    body = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in want]
    missing = want - {n.name for n in body}
    if missing:
        raise KeyError(f"{rel_path}: no top-level definition named {sorted(missing)}")
    mod = ast.Module(body=body, type_ignores=[])
    ns = monai_namespace()
    if extra_globals:
        ns.update(extra_globals)
    exec(compile(mod, path, "exec"), ns)   # line numbers in tracebacks are the reference's own
    return {n: ns[n] for n in want}


def sliding_window_two_heads():
    """The reference's two-head sliding_window_inference (trainer_CTUNet.py:417-581), unmodified."""
    return extract("trainer_CTUNet.py", ["sliding_window_inference", "_get_scan_interval"])["sliding_window_inference"]


def sliding_window_one_head():
    """The reference's one-head sliding_window_inference (trainer_CUNet.py:268-424), unmodified."""
    return extract("trainer_CUNet.py", ["sliding_window_inference", "_get_scan_interval"])["sliding_window_inference"]


# Shared sliding-window parity cases: (volume shape, roi, sw_batch, overlap, blend mode).  The same cases are run through
# the reference function (tests/test_reference_exec.py), stored as fixtures (tests/golden/make_golden.py) and through
# the CUDA blend on the GPU box (tests/test_sliding_window_gpu.py).
SW_CASES = [
    ((1, 1, 40, 40, 56), (32, 32, 32), 4, 0.5, "gaussian"),     # ragged last windows in every axis
    ((2, 1, 20, 33, 47), (16, 16, 16), 3, 0.25, "gaussian"),    # batch 2, window batches straddling the two images
    ((1, 1, 24, 10, 40), (16, 16, 16), 2, 0.7, "constant"),     # image smaller than the roi along y: padded, cropped
]


def sw_case_predictor(n_cls: int = 5):
    """Deterministic stand-in for the network with CTUNet's return structure (hybrid_CTUNet.py:857)."""
    def pred(w):
        base = torch.stack([torch.sin((c + 1.0) * w[:, 0]) + 0.1 * c for c in range(n_cls)], 1)
        return ((base, base * 0.5, base * 0.25), (torch.cos(base), base + 1.0))
    return pred
