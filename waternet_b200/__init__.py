"""waternet_b200 -- B200-native (sm_100a) implementation of the WaterNet hot path.

Drop-in for tnwei/waternet's preprocess (``waternet/data.py``) and gated-fusion
forward (``waternet/net.py``) behind the reference's own Python API.  The
arithmetic lives in ``libwaternet_b200.so`` (hand-written CUDA, C ABI in
``include/waternet_b200.h``); this package is the thin host side.
"""
from ._lib import MODE_BF16X3, MODE_DEFAULT, MODE_FP32_SIMT, WaterNetLibraryError  # noqa: F401

__all__ = ["MODE_BF16X3", "MODE_DEFAULT", "MODE_FP32_SIMT", "WaterNetLibraryError"]
