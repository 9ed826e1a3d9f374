"""Drop-in for the hot-path part of the reference's trainer_CTUNet.py: the two-head sliding-window inference
(trainer_CTUNet.py:417-581).  The training / validation loops of that file are host orchestration and stay
the reference's (out of scope)."""
from .sliding_window import get_scan_interval as _get_scan_interval
from .sliding_window import sliding_window_inference

__all__ = ["sliding_window_inference"]
