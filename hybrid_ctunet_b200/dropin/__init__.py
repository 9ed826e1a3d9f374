"""Switching the reference's own scripts to the sm_100a path without editing them.

    import hybrid_ctunet_b200.dropin as dropin
    dropin.install()            # `networks.hybrid_CTUNet`, `networks.resnet`, `networks.vit` now resolve to the drop-in modules
    import main_CTUNet          # the reference's script: builds hybrid_ctunet_b200's CTUNet, trains it with its own loop
    dropin.patch_trainers()     # trainer_*.sliding_window_inference -> the CUDA blend path (validation / test scripts)

`install()` puts this directory first on sys.path: it holds a `networks` package whose three modules re-export the drop-in
classes under the reference's module names (main_CTUNet.py:21, test_CTUNet.py:18, hybrid_CTUNet.py:21).  The trainers
(`trainer_CTUNet.py` etc.) stay the reference's files — their epoch loops are host orchestration — and only the
sliding-window function they define (trainer_CTUNet.py:417-557, trainer_CUNet.py:268-400) is rebound by `patch_trainers()`.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def install() -> str:
    """Make `import networks.*` resolve to the drop-in modules (idempotent).  Returns the path that was inserted."""
    if sys.path[:1] != [_HERE]:
        if _HERE in sys.path:
            sys.path.remove(_HERE)
        sys.path.insert(0, _HERE)
    for name in [n for n in sys.modules if n == "networks" or n.startswith("networks.")]:
        mod = sys.modules[name]
        if not (getattr(mod, "__file__", None) or "").startswith(_HERE):   # (a namespace package has __file__ None)
            del sys.modules[name]      # a previously imported reference `networks` package would win otherwise
    return _HERE


def patch_trainers() -> list:
    """Rebind `sliding_window_inference` of every already imported reference trainer module; returns their names."""
    from hybrid_ctunet_b200 import trainer_CTUNet, trainer_CUNet
    done = []
    for name, impl in (("trainer_CTUNet", trainer_CTUNet.sliding_window_inference),
                       ("trainer_CUNet", trainer_CUNet.sliding_window_inference),
                       ("trainer_TUNet", trainer_CUNet.sliding_window_inference)):
        mod = sys.modules.get(name)
        if mod is not None and hasattr(mod, "sliding_window_inference"):
            mod.sliding_window_inference = impl
            done.append(name)
    return done
