"""Import the UNMODIFIED reference networks from /root/reference through oracle/monai_stub.

Build-container only (the GPU box has no /root/reference): used by tests/golden/make_golden.py and by the CPU
tests that pin oracle/ against the reference.  Test infrastructure; never imported by the product package.
"""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("CTUNET_REFERENCE_ROOT", "/root/reference")
_STUB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "monai_stub")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "networks"))


def load():
    """Returns the reference modules (resnet, vit, hybrid_CTUNet)."""
    if not available():
        raise RuntimeError(f"{REFERENCE_ROOT} not present")
    try:
        import monai  # noqa: F401  (a real MONAI would be used if it existed)
    except ImportError:
        if _STUB not in sys.path:
            sys.path.insert(0, _STUB)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    sys.dont_write_bytecode = True  # /root/reference is read-only
    resnet = importlib.import_module("networks.resnet")
    vit = importlib.import_module("networks.vit")
    hyb = importlib.import_module("networks.hybrid_CTUNet")
    return resnet, vit, hyb
