"""argtypes/restype declarations for the non-GEMM entry points of include/ctunet_b200.h."""
import ctypes as C

_SIGS = {}


def sig(name, *argtypes):
    _SIGS[name] = argtypes


def declare(lib):
    for name, argtypes in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = list(argtypes)
        fn.restype = C.c_int
