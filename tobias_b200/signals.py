"""
Array-level mirrors of the reference's Cython signal functions (tobias/utils/signals.pyx), same names and semantics,
computed on the GPU through the C-ABI (K4 kernels).  float64 in, float64 out, like the reference.  A scratch context
on device 0 is created on first use (pass `ctx=` to reuse one).  There is no CPU fallback.
"""
import numpy as np

from . import lib

_ctx = None


def _default_ctx():
    global _ctx
    if _ctx is None:
        _ctx = lib.Context(0)
    return _ctx


def fast_rolling_math(arr, w, operation, ctx=None):
    """signals.pyx:102-177 -- "sum" / "mean" (the operations of the hot path; NaN in the flanks, C-float accumulator), "max" / "min" /
    "prod" with the reference's window clipping.  Any other operation returns the all-NaN array, as the reference does."""
    if operation not in ("sum", "mean", "max", "min", "prod"):
        return np.full(np.asarray(arr).shape[0], np.nan)
    return (ctx or _default_ctx()).rolling(np.asarray(arr, dtype=np.float64), int(w), operation)


def _score(kind, arr, flank_min, flank_max, fp_min, fp_max, ctx):
    ctx = ctx or _default_ctx()
    a = np.asarray(arr, dtype=np.float64)
    p = ctx.score_params(kind, flank=0, flank_min=flank_min, flank_max=flank_max, fp_min=fp_min, fp_max=fp_max, as_f64=True)
    return ctx.score_regions(a.astype(np.float32), [a.shape[0]], p)


def tobias_footprint_array(arr, flank_min, flank_max, fp_min, fp_max, ctx=None):
    """signals.pyx:184-230.  The reference truncates its input to C float (:191), so the float32 cast is part of the semantics."""
    return _score("footprint", arr, flank_min, flank_max, fp_min, fp_max, ctx)


def FOS_score(arr, flank_min, flank_max, fp_min, fp_max, ctx=None):
    """signals.pyx:237-282 (inputs are taken as float32, like the bigWig values ScoreBigwig feeds it)."""
    return _score("FOS", arr, flank_min, flank_max, fp_min, fp_max, ctx)
