"""Drop-in for the hot-path part of the reference's trainer_CUNet.py: the one-head sliding-window inference
(trainer_CUNet.py:268-424), also used for the TUNet member of the Hybrid ensemble (test_CTUNet_final.py:30,540)."""
from .sliding_window import get_scan_interval as _get_scan_interval
from .sliding_window import sliding_window_inference_one_head as sliding_window_inference

__all__ = ["sliding_window_inference"]
