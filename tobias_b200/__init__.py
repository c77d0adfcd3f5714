"""
tobias_b200 -- B200-native (sm_100a) implementation of the per-base signal hot path of TOBIAS
(ATACorrect bias estimation + bias correction, ScoreBigwig footprint scoring).

The compute lives in `csrc/` (hand-written CUDA behind the C-ABI of `include/tobias_b200.h`);
this package is the Python host that mirrors the reference's CLI / worker-function / Cython API
for that path and calls the C-ABI through ctypes.  There is no CPU fallback.
"""
__version__ = "0.1.0"
# version of the reference whose behaviour this drop-in mirrors (tobias/__init__.py:1)
REFERENCE_VERSION = "0.17.3"
