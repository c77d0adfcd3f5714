"""
The per-base steps of the two downstream tools that read the tracks this path writes, on the GPU through the C-ABI:

  score_bed_regions     `TOBIAS ScoreBed` (tobias/tools/score_bed.py:30-54, 118-131): one score per BED region and bigWig -- the value at the
                        start / middle / end of the region, or min / max / mean / sum over it; NaN -> 0; rounded to 5 decimals.
  aggregate_signal      `TOBIAS PlotAggregate` (tobias/tools/plot_aggregate.py:236-268): the signal of every site window (minus-strand sites read
                        back to front), outlier rows removed by the percentile of the row maxima, optional log transform, column means,
                        optional min-max normalisation and smoothing (fast_rolling_math "mean" on the edge-padded aggregate).

Reading the bigWig (io/bigwig.py) and the BED parsing stay on the host; the tools' plotting and argument handling are out of scope.
"""
import numpy as np

from . import signals


def _upload(ctx, bw, regions):
    sig, lens = [], []
    for chrom, start, end in regions:
        v = bw.values(chrom, start, end)
        sig.append(v)
        lens.append(v.shape[0])
    ctx.signal_upload(np.concatenate(sig) if sig else np.zeros(0, np.float32), lens)


def score_bed_regions(bw, regions, position="full", math="mean", null="0", ctx=None):
    """regions: [(chrom, start, end)]; returns the list ScoreBed would print for one bigWig: rounded floats, `null` for a contig the bigWig
    does not hold or an empty region."""
    ctx = ctx or signals._default_ctx()
    ok = [i for i, (c, s, e) in enumerate(regions) if c in bw.chroms and e > s]
    out = [null] * len(regions)
    if ok:
        _upload(ctx, bw, [regions[i] for i in ok])
        vals = ctx.signal_summary(position if position != "full" else math)
        for i, v in zip(ok, vals):
            out[i] = round(float(v), 5)
    return out


def aggregate_signal(bw, sites, width, remove_outliers=1.0, log_transform=False, normalize=False, smooth=1, ctx=None):
    """sites: [(chrom, start, end, strand)] already set to `width` around their centres.  Returns the aggregate (float64 [width])."""
    ctx = ctx or signals._default_ctx()
    if not sites:
        return np.zeros(width)
    _upload(ctx, bw, [s[:3] for s in sites])
    reverse = np.array([1 if (len(s) > 3 and s[3] == "-") else 0 for s in sites], dtype=np.uint8)
    max_values = ctx.signal_summary("max")
    upper = np.percentile(max_values, [100 * remove_outliers])[0]
    keep = (max_values <= upper).astype(np.uint8)
    agg = ctx.signal_aggregate(reverse, keep, log_transform)
    if normalize:
        lo, hi = agg.min(), agg.max()
        agg = (agg - lo) / (hi - lo) if hi > lo else np.zeros_like(agg)
    if smooth > 1:
        ext = np.pad(agg, smooth, "edge")
        agg = signals.fast_rolling_math(ext.astype("float64"), smooth, "mean", ctx=ctx)[smooth:-smooth]
    return agg
