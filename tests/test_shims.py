"""Array-level mirrors of the reference's Cython API, each against the compiled, unmodified reference (oracle/_ref) on
the same inputs: fast_rolling_math (every operation), SequenceMatrix.add_sequence / add_background (PWM and DWM, k-mers holding N,
negative amounts), ReadList.from_bam / split_strands / signal.  emu: kernel emulator (CPU); gpu: the CUDA build."""
import numpy as np
import pytest

from oracle import ref_loader
from tobias_b200 import ngs, sequences, signals
from tobias_b200.regions import OneRegion


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load()


def test_fast_rolling_math_every_operation(ctx, ref):
    rng = np.random.default_rng(3)
    a = rng.normal(0, 3, size=700)
    a[rng.random(700) < 0.2] = 0.0
    for w in (1, 6, 11, 100):
        for op in ("sum", "mean", "max", "min", "prod"):
            want = ref.signals.fast_rolling_math(a.copy(), w, op)
            got = signals.fast_rolling_math(a, w, op, ctx=ctx)
            if op == "prod":
                assert np.allclose(got, want, rtol=1e-13, atol=0, equal_nan=True), (w, op)      # 100 factors: product order differs by nothing, libm does not enter
            else:
                assert np.array_equal(got, want, equal_nan=True), (w, op)
    assert np.all(np.isnan(signals.fast_rolling_math(a, 5, "median", ctx=ctx)))              # unknown operation: all NaN, like the reference


@pytest.mark.parametrize("stype", ["PWM", "DWM"])
def test_add_sequence_and_background(ctx, ref, stype):
    rng = np.random.default_rng(5)
    L = 9
    mine = sequences.SequenceMatrix.create(L, stype)
    theirs = ref.sequences.SequenceMatrix.create(L, stype)
    kmers = rng.integers(0, 4, size=(60, L)).astype(np.int64)
    kmers[7, 3] = 4                                               # a k-mer holding N
    kmers[20, 0] = 4
    amounts = rng.integers(1, 11, size=60).astype(np.float64)
    if stype == "PWM":
        amounts[::5] *= -1.0                                      # verification counts: negative weights go to neg_counts (:81-84)
    for k, a in zip(kmers, amounts):
        theirs.add_sequence(k, float(a))
        theirs.add_background(k, float(abs(a)))
    for k, a in zip(kmers[:10], amounts[:10]):                    # one at a time, like the reference's callers ...
        mine.add_sequence(k, float(a), ctx=ctx)
        mine.add_background(k, float(abs(a)), ctx=ctx)
    mine.add_sequence(kmers[10:], amounts[10:], ctx=ctx)          # ... and as a batch
    mine.add_background(kmers[10:], np.abs(amounts[10:]), ctx=ctx)
    assert np.array_equal(mine.counts, np.asarray(theirs.counts))
    assert np.array_equal(mine.bg_counts, np.asarray(theirs.bg_counts))
    assert np.array_equal(mine.neg_counts, np.asarray(theirs.neg_counts))
    assert mine.no_bias == theirs.no_bias and mine.no_bg == theirs.no_bg
    mine.prepare_mat()
    theirs.prepare_mat()
    assert np.array_equal(mine.pssm, np.asarray(theirs.pssm), equal_nan=True)


def test_readlist_signal(ctx, ref, golden, tmp_path):
    npz = str(tmp_path / "reads.npz")
    golden.reads.save(npz)
    import pysam                                                 # the oracle's shim (installed by ref_loader) reads the same records
    bam = pysam.AlignmentFile(npz, "rb")
    for (chrom, a, b) in ((golden.names[0], 15000, 16200), (golden.names[1], 9000, 9500), (golden.names[0], 100, 400)):
        region = OneRegion([chrom, a, b])
        rl = ngs.ReadList().from_bam(golden.reads, region)
        theirs = ref.ngs.ReadList().from_bam(bam, ref.regions.OneRegion([chrom, a, b]))
        assert len(rl) == len(theirs)
        for r in theirs:
            r.get_cutsite([4, -5])
        mine_f, mine_r = rl.get_cutsites([4, -5]).split_strands()
        their_f, their_r = theirs.split_strands()
        inner = OneRegion([chrom, a + 40, b - 40])                # a region other than the fetched one: records outside still count
        for m, t in ((mine_f, their_f), (mine_r, their_r), (rl, theirs)):
            assert len(m) == len(t)
            for reg in (region, inner):
                want = t.signal(ref.regions.OneRegion([reg.chrom, reg.start, reg.end]))
                got = m.signal(reg, ctx=ctx)
                assert np.array_equal(got, want), (chrom, a, b, reg.start)
    assert rl.signal(OneRegion([golden.names[0], 50, 50]), ctx=ctx).shape == (0,)


def test_scorebed_and_plotaggregate_gathers(ctx, tmp_path):
    """the per-base steps of ScoreBed (tobias/tools/score_bed.py:30-54,118-131) and PlotAggregate (plot_aggregate.py:236-268)
    against their numpy statements in the reference, on a bigWig written by this package."""
    from tobias_b200 import gathers
    from tobias_b200.io import bigwig
    rng = np.random.default_rng(2)
    path = str(tmp_path / "t.bw")
    w = bigwig.BigWigWriter(path)
    w.add_header([("c1", 60000), ("c2", 5000)])
    pos = np.sort(rng.choice(60000, 30000, replace=False))
    val = rng.normal(0, 2, size=pos.size).astype(np.float32)
    w.add_entries("c1", pos, val)
    w.close()
    bw = bigwig.BigWigReader(path)
    full = np.zeros(60000, np.float32)
    full[pos] = val
    regions = [("c1", int(a), int(a) + int(n)) for a, n in zip(rng.integers(0, 59000, 40), rng.integers(1, 900, 40))] + [("c3", 5, 50), ("c2", 10, 60)]
    for position, math in (("start", None), ("mid", None), ("end", None), ("full", "min"), ("full", "max"), ("full", "mean"), ("full", "sum")):
        got = gathers.score_bed_regions(bw, regions, position, math or "mean", null="NA", ctx=ctx)
        for (c, s, e), g in zip(regions, got):
            if c == "c3":
                assert g == "NA"
                continue
            signal = np.nan_to_num(full[s:e] if c == "c1" else np.zeros(e - s, np.float32))
            want = {"start": lambda x: x[0], "mid": lambda x: x[int(len(x) / 2.0)], "end": lambda x: x[-1]}.get(position) or \
                {"min": np.min, "max": np.max, "mean": np.mean, "sum": np.sum}[math]
            assert abs(g - round(float(want(signal)), 5)) <= 2e-5 * max(1.0, abs(float(want(signal)))), (position, math, c, s, e)
    width = 120
    sites = [("c1", int(a), int(a) + width, "-" if k % 3 == 0 else "+") for k, a in enumerate(rng.integers(200, 59000, 300))]
    mat = np.array([np.nan_to_num(full[s:e])[::-1] if st == "-" else np.nan_to_num(full[s:e]) for (_c, s, e, st) in sites])
    for frac, logt, norm, smooth in ((1.0, False, False, 1), (0.9, True, False, 1), (1.0, False, True, 5)):
        got = gathers.aggregate_signal(bw, sites, width, frac, logt, norm, smooth, ctx=ctx)
        m = mat[np.max(mat, axis=1) <= np.percentile(np.max(mat, axis=1), [100 * frac])[0]]
        if logt:
            lg = np.log2(np.abs(m) + 1)
            lg[m < 0] *= -1
            m = lg
        want = np.nanmean(m.astype(np.float64), axis=0)
        if norm:
            want = (want - want.min()) / (want.max() - want.min())
        if smooth > 1:
            want = restate_roll(np.pad(want, smooth, "edge"), smooth)[smooth:-smooth]
        assert np.allclose(got, want, rtol=1e-6, atol=1e-7), (frac, logt, norm, smooth)
    bw.close()


def restate_roll(a, w):
    from oracle import restate
    return restate.fast_rolling_math(a.astype(np.float64), w, "mean")
