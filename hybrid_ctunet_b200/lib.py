"""ctypes binding of include/ctunet_b200.h.

The library is loaded lazily; if it is missing, cannot be loaded, or the device is not an sm_100 GPU, every
operator raises — there is no CPU or eager-PyTorch fallback behind these calls.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_LIB = None
_DEVICE_OK = False


class CtuError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("w", C.c_void_p), ("out", C.c_void_p), ("bias", C.c_void_p),
        ("residual", C.c_void_p), ("stats", C.c_void_p),
        ("a_c", C.c_int32), ("lda", C.c_int32),
        ("d1", C.c_int32), ("d2", C.c_int32), ("d3", C.c_int32), ("d4", C.c_int32),
        ("b1", C.c_int32), ("b2", C.c_int32), ("b3", C.c_int32),
        ("k1", C.c_int32), ("k2", C.c_int32), ("k3", C.c_int32),
        ("n_pad", C.c_int32), ("n_real", C.c_int32), ("k_total", C.c_int32),
        ("block_n", C.c_int32),
        ("out_mode", C.c_int32), ("ldc", C.c_int32),
        ("act", C.c_int32),
        ("res_mode", C.c_int32), ("ldr", C.c_int32),
        ("convt_cout", C.c_int32), ("u1", C.c_int32), ("u2", C.c_int32), ("u3", C.c_int32),
        ("stats_ld", C.c_int32), ("out_col0", C.c_int32), ("a_c_live", C.c_int32),
        ("w_x3", C.c_void_p),
    ]


class WgradDesc(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("dy", C.c_void_p), ("dw", C.c_void_p),
        ("x_c", C.c_int32), ("ldx", C.c_int32), ("n", C.c_int32), ("ldy", C.c_int32), ("ldw", C.c_int32),
        ("d1", C.c_int32), ("d2", C.c_int32), ("d3", C.c_int32), ("d4", C.c_int32),
        ("b1", C.c_int32), ("b2", C.c_int32), ("b3", C.c_int32),
        ("k1", C.c_int32), ("k2", C.c_int32), ("k3", C.c_int32),
        ("block_n", C.c_int32),
    ]


LOSS_MAX_HEADS = 8


class LossHeads(C.Structure):
    """ctu_loss_heads of include/ctunet_b200.h (field for field)."""
    _fields_ = [
        ("n_heads", C.c_int32),
        ("B", C.c_int32 * LOSS_MAX_HEADS), ("C", C.c_int32 * LOSS_MAX_HEADS),
        ("S", C.c_int64 * LOSS_MAX_HEADS),
        ("sums_off", C.c_int64 * LOSS_MAX_HEADS), ("coef_off", C.c_int64 * LOSS_MAX_HEADS),
        ("weight", C.c_double * LOSS_MAX_HEADS),
        ("lambda_dice", C.c_double), ("lambda_ce", C.c_double), ("smooth_nr", C.c_double), ("smooth_dr", C.c_double),
    ]


class InvertGeom(C.Structure):
    """ctu_invert_geom of include/ctunet_b200.h (field for field)."""
    _fields_ = [
        ("m", C.c_double * 12),
        ("out_size", C.c_int32 * 3), ("pad_size", C.c_int32 * 3), ("crop_start", C.c_int32 * 3), ("pred_size", C.c_int32 * 3),
        ("mode", C.c_int32),
    ]


def lib_path() -> str:
    return _build.LIB_PATH


def load(build_if_missing: bool = True):
    """Load libctunet_b200.so (building it with nvcc first if it is absent or stale)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("CTU_LIB") or _build.LIB_PATH  # CTU_LIB: another build of the same library (A/B timing)
    if "CTU_LIB" in os.environ:
        build_if_missing = False
    if not os.path.exists(path) or (build_if_missing and not _build.is_fresh() and _have_nvcc()):
        if not build_if_missing:
            raise CtuError(f"{path} is missing; run `python -m hybrid_ctunet_b200.build`")
        path = _build.build()
    try:
        lib = C.CDLL(path)
    except OSError as e:  # pragma: no cover
        raise CtuError(f"cannot load {path}: {e}") from e
    _declare(lib)
    _LIB = lib
    return lib


def _have_nvcc() -> bool:
    try:
        _build._nvcc()
        return True
    except RuntimeError:
        return False


def _declare(lib):
    lib.ctu_umma_gemm.argtypes = [C.POINTER(GemmDesc), C.c_void_p]
    lib.ctu_umma_gemm.restype = C.c_int
    lib.ctu_umma_wgrad.argtypes = [C.POINTER(WgradDesc), C.c_void_p]
    lib.ctu_umma_wgrad.restype = C.c_int
    lib.ctu_pack_item_tasks.argtypes = [C.c_int] * 7
    lib.ctu_pack_item_tasks.restype = C.c_longlong
    lib.ctu_set_persistent_sm_limit.argtypes = [C.c_int]
    lib.ctu_set_persistent_sm_limit.restype = None
    lib.ctu_launch_count.argtypes = []
    lib.ctu_launch_count.restype = C.c_int64
    lib.ctu_device_ok.argtypes = []
    lib.ctu_device_ok.restype = C.c_int
    lib.ctu_version.argtypes = []
    lib.ctu_version.restype = C.c_char_p
    from . import _abi
    _abi.declare(lib)


def check(rc: int, what: str):
    if rc == 0:
        return
    if rc < 0:
        names = {-1: "bad argument", -2: "unsupported configuration", -3: "driver entry point / TMA encode failed"}
        raise CtuError(f"{what}: {names.get(rc, rc)}")
    raise CtuError(f"{what}: CUDA error {rc}")


def launch_count() -> int:
    return int(load().ctu_launch_count())


def require_device():
    """Fail loudly unless the CUDA library is loaded and the current device is a B200-class (sm_100) GPU."""
    global _DEVICE_OK
    lib = load()
    if not _DEVICE_OK:  # cudaGetDeviceProperties is slow: query once
        if not lib.ctu_device_ok():
            raise CtuError("ctunet_b200 needs an sm_100 (B200) GPU; there is no CPU fallback")
        _DEVICE_OK = True
    return lib
