"""Size-independent properties of the CUDA hot path on a benchmark-shaped workload (a 12 Mb contig, 3000 peaks, 1.2M records --
too large for the oracle in CI time, the same generator and code path as bench.py's hg38 configuration).  GPU only.

  * counts are conserved:  sum of the pileups == cut sites inside the regions (independent numpy count on the host);
    every (m, n) plane of a DWM count tensor sums to the number of weighted k-mers; PWM columns likewise
  * the uncorrected track is pileup * correction_factor exactly; expected >= 0; corrected = uncorrected - expected (x scale on
    the positive part, scale >= 1); the bias track is finite
  * the whole path is deterministic (two runs: identical tracks) and independent of how the regions are batched
    (one call vs two calls over halves: identical per-region results -- LM window grids are region-relative)
  * the K2 slab is left clean (second call: identical integers); chunked slab == unchunked
  * footprint scores are >= 0, zero where the signal is zero everywhere, and invariant to appending regions
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def work():
    import torch
    from tobias_b200 import lib, synth_torch
    from tobias_b200.pipeline import prepare_regions
    from tobias_b200.regions import OneRegion, RegionList, to_arrays
    ds = synth_torch.TorchDataset(4242, {"chr1": 12_000_000}, 3000, 1_200_000, device="cuda" if torch.cuda.is_available() else "cpu")
    chrom_info = dict(zip(ds.names, ds.lengths))
    peaks = RegionList([OneRegion(list(p)) for p in ds.peaks])
    regs = prepare_regions(peaks, chrom_info, ds.names, extend=100, k_flank=12, window=100)
    index = {n: i for i, n in enumerate(ds.names)}
    arrays = {k: to_arrays(sorted(regs[k], key=lambda r: (index[r.chrom], r.start, r.end)), index)
              for k in ("peak_regions", "nonpeak_regions", "input_regions", "output_regions")}
    ctx = lib.Context(0)
    assert not ctx.device_info()["is_emulator"]
    ctx.genome_load(ds.genome)
    ctx.reads_upload_columns(ds.reads)
    yield ctx, ds, arrays
    ctx.close()


def _cut_sites(reads, shift=(4, -5)):
    """0-based pileup index of every usable record (ngs.pyx:91-106), and its strand."""
    rev = reads.is_reverse
    clip = reads.clip5()
    pos = np.where(rev, reads.ref_end.astype(np.int64) + clip + shift[1], reads.ref_start.astype(np.int64) - clip + shift[0])
    return pos, rev, reads.usable()


def test_pileup_conserves_cut_sites(work):
    ctx, ds, arrays = work
    c, s, e = arrays["output_regions"]
    fwd, rev = ctx.pileup(c, s, e)
    pos, is_rev, ok = _cut_sites(ds.reads)
    # regions are disjoint: a cut site lies in at most one.  ReadList.signal + the filter at atacorrect_functions.py:237:
    # the alignment overlaps the region and start < cutsite (1-based) < end  <=>  s <= pos <= e - 2
    rs, re = ds.reads.ref_start.astype(np.int64), ds.reads.ref_end.astype(np.int64)
    j = np.clip(np.searchsorted(s, pos, side="right") - 1, 0, len(s) - 1)
    inside = ok & (pos >= s[j]) & (pos <= e[j] - 2) & (rs < e[j]) & (re > s[j])
    assert int(fwd.sum()) == int((inside & ~is_rev).sum())
    assert int(rev.sum()) == int((inside & is_rev).sum())
    assert fwd.sum() > 10000


def test_count_tensors_are_consistent(work):
    ctx, ds, arrays = work
    regs = arrays["input_regions"]
    counts, scalars = ctx.bias_counts(*regs, is_dwm=True)
    again, scalars2 = ctx.bias_counts(*regs, is_dwm=True)
    assert np.array_equal(counts, again) and np.array_equal(scalars, scalars2)          # slab left clean, deterministic integers
    for si in range(2):
        for wi in range(2):
            t = counts[si, wi][:4, :4]
            planes = t.sum(axis=(0, 1))                                               # [L, L]: weighted k-mers per (m, n)
            assert np.all(planes == planes[0, 0]) and planes[0, 0] == scalars[1 + 2 * si + wi]
            assert np.array_equal(t, t.transpose(1, 0, 3, 2))                         # symmetric in (Sm, m) <-> (Sn, n)
            diag = np.stack([t[a, a][np.arange(25), np.arange(25)] for a in range(4)])   # m == n: base counts per position
            off = t[:, :, np.arange(25), np.arange(25)]
            assert np.all(off[~np.eye(4, dtype=bool)] == 0) and np.all(diag.sum(axis=0) == planes[0, 0])
    pwm, pscal = ctx.bias_counts(*regs, is_dwm=False)
    assert pscal[0] == scalars[0]
    for si in range(2):
        assert np.array_equal(pwm[si, 0][:4].sum(axis=0), np.full(25, scalars[1 + 2 * si]))      # same k-mers as the DWM run
        assert np.array_equal(pwm[si, 0][:4], np.stack([counts[si, 0][a, a][np.arange(25), np.arange(25)] for a in range(4)]))


def test_slab_chunking_is_invisible(work, monkeypatch):
    ctx, ds, arrays = work
    regs = arrays["input_regions"]
    counts, scalars = ctx.bias_counts(*regs, is_dwm=True)
    ctx.set_option("k2_tile", 4096)
    chunked, scalars2 = ctx.bias_counts(*regs, is_dwm=True)
    ctx.set_option("k2_tile", -1)
    assert np.array_equal(counts, chunked) and np.array_equal(scalars, scalars2)


def test_count_kernels_agree(work, monkeypatch):
    """The three DWM count kernels -- joint histograms in shared memory (default), joint histograms in L2, one shared-memory atomic
    per position pair -- produce the same integers on the benchmark-shaped workload (1.2M records, cap path included)."""
    ctx, ds, arrays = work
    regs = arrays["input_regions"]
    counts, scalars = ctx.bias_counts(*regs, is_dwm=True)
    ctx.set_option("k2_no_smem_hist", 1)
    l2, l2_sc = ctx.bias_counts(*regs, is_dwm=True)
    ctx.set_option("k2_no_hist", 1)
    pairs, pairs_sc = ctx.bias_counts(*regs, is_dwm=True)
    ctx.set_option("k2_no_smem_hist", -1)
    ctx.set_option("k2_no_hist", -1)
    assert np.array_equal(counts, l2) and np.array_equal(scalars, l2_sc)
    assert np.array_equal(counts, pairs) and np.array_equal(scalars, pairs_sc)
    assert scalars[0] > 100000


def _model(ctx, arrays):
    from tobias_b200.pipeline import AtacBias, STRANDS
    counts, scalars = ctx.bias_counts(*arrays["input_regions"], is_dwm=True)
    bias = AtacBias(25, "DWM")
    for si, st in enumerate(STRANDS):
        m = bias.bias[st]
        m.counts, m.bg_counts = counts[si, 0].astype(np.float64), counts[si, 1].astype(np.float64)
        m.prepare_mat()
        ctx.set_bias_model(si, True, 25, m.bias_pwm_log, m.bg_pwm_log, m.bias_dwm_log, m.bg_dwm_log)


def test_correction_identities_determinism_and_batching(work):
    ctx, ds, arrays = work
    _model(ctx, arrays)
    c, s, e = arrays["output_regions"]
    cf = 1.7
    ctx.correct_run(c, s, e, correction_factor=cf)
    a = ctx.correct_download(split_strands=True, as_f64=True, prepost=False)
    # bias_correction piles up over the region extended by k_flank + window / 2 = 62 and trims afterwards (:220-243, :381-383)
    xf, xr = ctx.pileup(c, s - 62, e + 62)
    keep = np.concatenate([np.arange(o + 62, o + 62 + n) for o, n in zip(np.concatenate([[0], np.cumsum(e - s + 124)[:-1]]), e - s)])
    fwd, rev = xf[keep], xr[keep]
    for si, pile in enumerate((fwd, rev)):
        unc, exp, cor, bias = a["uncorrected"][si], a["expected"][si], a["corrected"][si], a["bias"][si]
        assert np.array_equal(unc, pile.astype(np.float64) * cf)
        assert np.all(exp >= 0) and np.all(np.isfinite(bias)) and np.all(np.isfinite(cor))
        raw = unc - exp
        neg = raw <= 0
        assert np.array_equal(cor[neg], raw[neg])                       # only positive values are rescaled (:345-350)
        pos = raw > 0
        assert np.all(cor[pos] >= raw[pos] * (1 - 1e-15))               # scale >= 1
    ctx.correct_run(c, s, e, correction_factor=cf)
    b = ctx.correct_download(split_strands=True, as_f64=True, prepost=False)
    for t in a:
        for si in range(2):
            assert np.array_equal(a[t][si], b[t][si]), t                # deterministic (no order-dependent arithmetic in the tracks)
    # the same regions in two calls (different list compaction, tiles, launch shapes): identical per-region results
    h = len(c) // 2
    off = int((e[:h] - s[:h]).sum())
    ctx.correct_run(c[:h], s[:h], e[:h], correction_factor=cf)
    p1 = ctx.correct_download(split_strands=True, as_f64=True, prepost=False)
    ctx.correct_run(c[h:], s[h:], e[h:], correction_factor=cf)
    p2 = ctx.correct_download(split_strands=True, as_f64=True, prepost=False)
    for t in a:
        for si in range(2):
            assert np.array_equal(a[t][si][:off], p1[t][si]) and np.array_equal(a[t][si][off:], p2[t][si]), t
    trace = ctx.correct_fit_trace()
    assert set(np.unique(trace[:, :, 0])) <= {0.0, 1.0, 2.0}


def test_fast_and_exact_fits_agree(work):
    """lm_exact_order = 0 (one-pass outer iterations, DWM default) against 1 (MINPACK-exact arithmetic) on ~200k window fits:
    same fit / fallback / empty decisions except for a handful of windows, tracks within the DWM-mode tolerance."""
    ctx, ds, arrays = work
    _model(ctx, arrays)
    c, s, e = arrays["output_regions"]
    ctx.correct_run(c, s, e, correction_factor=1.0, lm_exact_order=1)
    a = ctx.correct_download(split_strands=True, as_f64=True, prepost=False)
    ta = ctx.correct_fit_trace()
    ctx.correct_run(c, s, e, correction_factor=1.0, lm_exact_order=0)
    b = ctx.correct_download(split_strands=True, as_f64=True, prepost=False)
    tb = ctx.correct_fit_trace()
    assert (ta[:, :, 0] != tb[:, :, 0]).mean() < 1e-3                          # decisions
    assert abs(ta[:, :, 5].mean() - tb[:, :, 5].mean()) < 0.05                 # same amount of work (nfev)
    for t in ("expected", "corrected"):
        for si in range(2):
            d = np.abs(a[t][si] - b[t][si])
            ok = d <= 1e-5 * np.abs(a[t][si]) + 2e-6
            assert ok.mean() >= 0.999, (t, si, ok.mean(), d.max())
    for si in range(2):
        assert np.array_equal(a["uncorrected"][si], b["uncorrected"][si]) and np.array_equal(a["bias"][si], b["bias"][si])


def test_footprints_properties(work):
    ctx, ds, arrays = work
    _model(ctx, arrays)
    c, s, e = arrays["output_regions"]
    ctx.correct_run(c, s, e, correction_factor=1.0)
    ctx.signal_from_corrected(30, e - s)
    p = ctx.score_params("footprint", flank=30, as_f64=False)
    ctx.score_run(p)
    fp = ctx.score_download()
    assert fp.shape[0] == int((e - s).sum()) and np.all(fp >= 0) and np.all(np.isfinite(fp)) and fp.max() > 0
    # appending regions does not change the scores of the first ones; an all-zero region scores zero
    both = ctx.correct_download(split_strands=False, as_f64=False, tracks=("corrected",), prepost=False)["corrected"]
    lens = (e - s)[:50]
    n = int(lens.sum())
    sig = np.concatenate([np.pad(x, 30) for x in np.split(both[:n], np.cumsum(lens)[:-1])]).astype(np.float32)
    part = ctx.score_regions(sig, lens + 60, p)
    assert np.array_equal(part, fp[:n])
    zero = ctx.score_regions(np.zeros(5000, dtype=np.float32), [5000], ctx.score_params("footprint", flank=30, as_f64=False))
    assert not zero.any()


def test_whole_contig_footprint_scan_matches_the_oracle_on_slices():
    """BASELINE.json configs[3]: ScoreBigwig footprint scan, fp widths 20-50, flank 100, over a whole-contig corrected track
    (one region of 120 Mb: ~300k tiles; dense float32 track, ~70 % exact zeros).  The score of a position depends on the
    signal within flank_max + fp_max of it only, so random slices (incl. both contig ends) are checked against the oracle's
    restatement of tobias_footprint_array run on the slice + halo; plus: an all-zero stretch scores zero."""
    from oracle import restate
    from tobias_b200 import lib
    n = 120_000_000
    rng = np.random.default_rng(77)
    sig = np.zeros(n, dtype=np.float32)
    nz = rng.random(n) >= 0.7
    sig[nz] = np.round(rng.normal(0.0, 0.5, size=int(nz.sum())), 5).astype(np.float32)
    sig[50_000_000:50_100_000] = 0.0
    args = (100, 100, 20, 50)
    halo = args[1] + args[3]
    with lib.Context(0) as ctx:
        p = ctx.score_params("footprint", flank=0, flank_min=args[0], flank_max=args[1], fp_min=args[2], fp_max=args[3], as_f64=False)
        got = ctx.score_regions(sig, [n], p)
        ms = ctx.last_kernel_ms()
    assert got.shape == (n,) and np.all(np.isfinite(got)) and np.all(got >= 0)
    assert not got[50_000_000 + halo:50_100_000 - halo].any()
    starts = [0, n - 3000] + [int(x) for x in rng.integers(halo, n - 3000 - halo, size=10)] + [50_000_000 - 1500]
    for a in starts:
        lo, hi = max(0, a - 2 * halo), min(n, a + 3000 + 2 * halo)
        ref = restate.tobias_footprint_array(sig[lo:hi].astype(np.float64), *args)
        # positions whose windows all lie inside the slice (or at a true contig end) are complete in the slice's result
        a0 = a if lo > 0 else 0
        b0 = a + 3000 if hi < n else n
        r, g = ref[a0 - lo:b0 - lo], got[a0:b0].astype(np.float64)
        assert np.array_equal(g == 0, r == 0), a
        m = r != 0
        assert np.all(np.abs(g[m] - r[m]) <= 1e-5 * np.abs(r[m])), a
    print("whole-contig footprint scan: %.1f ms for %d bp (%.3g bp/s)" % (ms, n, n / (ms * 1e-3)))
