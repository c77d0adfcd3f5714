"""Parity of the CUDA hot path (through the C-ABI) against the golden fixtures produced by the reference and against the
CPU oracle on the same inputs.  Every test runs twice: `emu` (kernel-logic emulator, CPU, `-m "not gpu"`) and `gpu`
(the real sm_100a build on the B200, `-m gpu`).

Tolerances (stated per the north star):
  * pileups, read counts, count tensors                     bit-exact (integers)
  * bias track (DWM / PWM score)                            |d| <= 1e-9 * max(1, |ref|)   (fp64; only exp2/log2 and sum order differ)
  * footprint / rolling scores                              rtol 1e-5 (footprint: fp64 prefix differences vs sequential fp32 sums);
                                                            rolling sums/means reproduce the fp32 accumulator exactly
  * expected / corrected, PWM mode                          bit-exact against the reference (identical bias input, exact-order LM)
  * expected / corrected, DWM mode                          rtol 1e-5 + atol 2e-6 on >= 99.9 % of positions, atol 2e-5 everywhere:
        the window fits differentiate by forward differences (h ~ 1.5e-8), which amplifies the ~1e-13 difference of the
        bias input (CUDA vs glibc exp2/log2) to ~1e-8 in the fitted parameters -- the reference shows the same scatter
        against itself under such a perturbation (tests/test_lm_exact.py::test_reference_sensitivity)."""
import os

import numpy as np
import pytest

from oracle import restate
from tobias_b200 import lib

TRACKS = ("uncorrected", "bias", "expected", "corrected")


def _cols(a):
    a = np.asarray(a)
    return a[:, 0], a[:, 1], a[:, 2]


def test_genome_roundtrip(loaded_ctx, golden):
    for c, seq in enumerate(golden.genome):
        assert np.array_equal(loaded_ctx.genome_fetch(c, 0, len(seq)), restate.encode(seq))
    assert np.array_equal(loaded_ctx.genome_fetch(0, 777, 1001), restate.encode(golden.genome[0][777:1778]))


def test_count_reads_bit_exact(loaded_ctx, golden):
    g = golden.atac
    assert loaded_ctx.count_reads(*_cols(g["nonpeak"])) == int(g["count_nonpeak"])
    assert loaded_ctx.count_reads(*_cols(g["peakreg"])) == int(g["count_peak"])


def test_pileup_bit_exact(loaded_ctx, golden):
    c, s, e = _cols(golden.atac["outreg"])
    fwd, rev = loaded_ctx.pileup(c, s - 62, e + 62)
    off = 0
    for ci, a, b in zip(c, s - 62, e + 62):
        f0, r0 = restate.pileup(golden.reads, int(ci), int(a), int(b))
        assert np.array_equal(fwd[off:off + b - a], f0) and np.array_equal(rev[off:off + b - a], r0)
        off += b - a
    assert fwd.sum() > 0 and rev.sum() > 0


def test_pileup_empty_and_edges(loaded_ctx, golden):
    # a region inside an N block holds no reads; a 1-bp region; the first region of a contig
    fwd, rev = loaded_ctx.pileup([0, 0, 1], [10, 5000, 0], [11, 5001, 300])
    assert fwd.shape == (302,) and fwd[:1].sum() == 0
    f0, r0 = restate.pileup(golden.reads, 1, 0, 300)
    assert np.array_equal(fwd[2:], f0) and np.array_equal(rev[2:], r0)
    with pytest.raises(lib.TobiasB200Error):
        loaded_ctx.pileup([0, 0], [500, 100], [600, 200])          # unsorted
    with pytest.raises(lib.TobiasB200Error):
        loaded_ctx.pileup([0], [100], [10 ** 9])                   # outside the contig


@pytest.mark.parametrize("mode", ["DWM", "PWM"])
def test_bias_counts_bit_exact(loaded_ctx, golden, mode):
    g = golden.atac
    counts, scalars = loaded_ctx.bias_counts(*_cols(g["nonpeak"]), is_dwm=(mode == "DWM"))
    assert scalars[0] == int(g[mode + "_no_reads"])
    for si, st in enumerate(("forward", "reverse")):
        assert np.array_equal(counts[si, 0], g["%s_%s_counts" % (mode, st)]), st
        assert np.array_equal(counts[si, 1], g["%s_%s_bg_counts" % (mode, st)]), st
        assert [scalars[1 + 2 * si], scalars[2 + 2 * si]] == list(g["%s_%s_no" % (mode, st)])


def test_bias_counts_other_parameters(loaded_ctx, golden):
    regs = [tuple(int(x) for x in r) for r in golden.atac["nonpeak"]]
    ref = restate.bias_estimation(golden.reads, golden.genome, regs, golden.lengths, k_flank=5, bg_shift=37, cap=3)
    counts, scalars = loaded_ctx.bias_counts(*_cols(golden.atac["nonpeak"]), k_flank=5, bg_shift=37, cap=3)
    for si, st in enumerate(("forward", "reverse")):
        assert np.array_equal(counts[si, 0], ref.bias[st].counts) and np.array_equal(counts[si, 1], ref.bias[st].bg_counts)
    assert scalars[0] == ref.no_reads


def test_bias_counts_in_several_slab_tiles(loaded_ctx, golden, monkeypatch):
    """A region longer than the shared-memory multiplicity slab is walked tile by tile; a tiny tile forces many tiles per region (and,
    with more distinct sites than the block's list holds, the dense-scan path) and a tiny site list the growth path -- same integers."""
    g = golden.atac
    loaded_ctx.set_option("k2_tile", 1024)
    counts, scalars = loaded_ctx.bias_counts(*_cols(g["nonpeak"]), is_dwm=True)
    loaded_ctx.set_option("k2_tile", -1)
    assert scalars[0] == int(g["DWM_no_reads"])
    for si, st in enumerate(("forward", "reverse")):
        assert np.array_equal(counts[si, 0], g["DWM_%s_counts" % st]) and np.array_equal(counts[si, 1], g["DWM_%s_bg_counts" % st])
    counts2, scalars2 = loaded_ctx.bias_counts(*_cols(g["nonpeak"]), is_dwm=True)           # slab left clean by the tiled run
    assert np.array_equal(counts, counts2) and np.array_equal(scalars, scalars2)
    loaded_ctx.set_option("k2_site_cap", 100)                                            # site list too small: regrown, chunk redone
    counts3, scalars3 = loaded_ctx.bias_counts(*_cols(g["nonpeak"]), is_dwm=True)
    loaded_ctx.set_option("k2_site_cap", -1)
    assert np.array_equal(counts, counts3) and np.array_equal(scalars, scalars3)


def test_bias_counts_histogram_forms(loaded_ctx, golden, monkeypatch):
    """DWM counts come from joint chunk histograms (one atomic per k-mer and chunk pair) folded into position pairs: the
    shared-memory form (chunks of 2 bases, the default; also chunks of 1), the L2 form (every chunking -- including a last chunk
    of one base and two chunks only) and the shared-memory-atomic kernel (option k2_no_hist, models of a single chunk) give the
    same integers as the reference."""
    regs = [tuple(int(x) for x in r) for r in golden.atac["nonpeak"]]
    emu = loaded_ctx.device_info()["is_emulator"]
    for k_flank, chunkings, smem_chunks in ((12, (3, 4) if emu else (3, 4, 5), (2,)), (5, (1, 2, 5), (1, 2)), (2, (3, 4, 5), (1, 2)), (1, (2,), (2,))):
        kw = dict(k_flank=k_flank, bg_shift=37, cap=3)
        ref = restate.bias_estimation(golden.reads, golden.genome, regs, golden.lengths, **kw)
        loaded_ctx.set_option("k2_no_hist", 1)
        base, base_sc = loaded_ctx.bias_counts(*_cols(golden.atac["nonpeak"]), **kw)
        loaded_ctx.set_option("k2_no_hist", -1)
        for si, st in enumerate(("forward", "reverse")):
            assert np.array_equal(base[si, 0], ref.bias[st].counts) and np.array_equal(base[si, 1], ref.bias[st].bg_counts)
        assert base_sc[0] == ref.no_reads
        for ch in smem_chunks:
            loaded_ctx.set_option("k2_smem_chunk", ch)
            got, got_sc = loaded_ctx.bias_counts(*_cols(golden.atac["nonpeak"]), **kw)
            loaded_ctx.set_option("k2_smem_chunk", -1)
            assert np.array_equal(got, base) and np.array_equal(got_sc, base_sc), ("smem", k_flank, ch)
        loaded_ctx.set_option("k2_no_smem_hist", 1)
        for ch in chunkings:
            loaded_ctx.set_option("k2_chunk", ch)
            got, got_sc = loaded_ctx.bias_counts(*_cols(golden.atac["nonpeak"]), **kw)
            loaded_ctx.set_option("k2_chunk", -1)
            assert np.array_equal(got, base) and np.array_equal(got_sc, base_sc), ("l2", k_flank, ch)
        loaded_ctx.set_option("k2_no_smem_hist", -1)


def test_asynchronous_genome_upload(loaded_ctx, golden):
    """tb_genome_upload_async + device-side fences: counts taken while the genome is still in flight (per-chunk waits in
    tb_bias_counts, several chunks) equal the counts over the synchronously loaded genome; the host buffers are released by
    genome_wait.  The fixture's genome is restored afterwards."""
    kw = dict(k_flank=12, bg_shift=100, cap=10)
    base, base_sc = loaded_ctx.bias_counts(*_cols(golden.atac["nonpeak"]), **kw)
    codes = [loaded_ctx.genome_fetch(i, 0, n) for i, n in enumerate(golden.lengths)]
    try:
        loaded_ctx.genome_create(golden.lengths)
        bufs = []
        for i, g in enumerate(golden.genome):
            b = loaded_ctx.pinned_empty(g.shape, np.uint8)
            b[...] = g
            bufs.append(b)
            loaded_ctx.genome_upload_async(i, b)
        got, got_sc = loaded_ctx.bias_counts(*_cols(golden.atac["nonpeak"]), **kw)
        loaded_ctx.genome_wait()
        assert np.array_equal(got, base) and np.array_equal(got_sc, base_sc)
        for i, n in enumerate(golden.lengths):
            assert np.array_equal(loaded_ctx.genome_fetch(i, 0, n), codes[i])
    finally:
        loaded_ctx.genome_load(golden.genome)


def test_packed_genome_upload(loaded_ctx, golden, tmp_path):
    """tb_genome_upload_packed: contigs packed on the host (hostpack.c; cached next to the FASTA as .tb2bit) land in the same device
    layout as ASCII contigs packed on the device -- synchronous and on the copy stream.  The fixture's genome is restored."""
    from tobias_b200.io.fasta import FastaFile, PackedGenome, pack_ascii, write_fasta
    codes = [loaded_ctx.genome_fetch(i, 0, n) for i, n in enumerate(golden.lengths)]
    fa = str(tmp_path / "g.fa")
    write_fasta(fa, zip(golden.names, golden.genome))
    pg = PackedGenome.from_fasta(FastaFile(fa))
    assert os.path.exists(fa + ".tb2bit") and pg.references == golden.names and pg.lengths == golden.lengths
    again = PackedGenome.from_fasta(FastaFile(fa), names=golden.names[::-1])          # from the cache, another order
    assert again.references == golden.names[::-1] and np.array_equal(again.seq2[0], pg.seq2[-1])
    try:
        for wait in (True, False):
            loaded_ctx.genome_create(golden.lengths)
            for i, n in enumerate(golden.lengths):
                sq, nm = (pg.seq2[i], pg.nmask[i]) if wait else pack_ascii(golden.genome[i])
                loaded_ctx.genome_upload_packed(i, sq, nm, n, wait=wait)
            if not wait:
                loaded_ctx.genome_wait()
            for i, n in enumerate(golden.lengths):
                assert np.array_equal(loaded_ctx.genome_fetch(i, 0, n), codes[i]), (wait, i)
        with pytest.raises(lib.TobiasB200Error):
            loaded_ctx.genome_upload_packed(0, pg.seq2[1], pg.nmask[1], golden.lengths[1])     # wrong contig length
    finally:
        loaded_ctx.genome_load(golden.genome)


def test_compact_inputs(loaded_ctx, golden):
    """tb_genome_upload_packed_runs (N mask as runs) and tb_reads_upload_compact (9-byte records) leave the device in the state of the
    plain uploads: same base codes, same pileup and read counts.  The fixture's genome and reads are restored."""
    from tobias_b200.io.fasta import pack_ascii
    rc = golden.reads
    codes = [loaded_ctx.genome_fetch(i, 0, n) for i, n in enumerate(golden.lengths)]
    regs = golden.atac["outreg"]
    c, s, e = _cols(regs)
    base_pile = loaded_ctx.pileup(c, s, e)
    base_count = loaded_ctx.count_reads(c, s, e)
    flags = (rc.is_reverse.astype(np.uint8) * lib.TB_READ_REVERSE | rc.five_prime_is_match().astype(np.uint8) * lib.TB_READ_5P_MATCH
             | (~rc.usable()).astype(np.uint8) * lib.TB_READ_SKIP)
    try:
        for wait in (True, False):
            loaded_ctx.genome_create(golden.lengths)
            for i, n in enumerate(golden.lengths):
                sq, nm = pack_ascii(golden.genome[i])
                rs, re_ = lib.mask_runs(nm, n)
                assert np.array_equal(np.flatnonzero(codes[i] == 4), np.concatenate([np.arange(a, b) for a, b in zip(rs, re_)] + [np.zeros(0, np.int64)]))
                loaded_ctx.genome_upload_packed_runs(i, sq, rs, re_, n, wait=wait)
            if not wait:
                loaded_ctx.genome_wait()
            for i, n in enumerate(golden.lengths):
                assert np.array_equal(loaded_ctx.genome_fetch(i, 0, n), codes[i]), (wait, i)
        comp = lib.compact_read_columns(rc.chrom_id, rc.ref_start, rc.ref_end, rc.clip5(), flags, len(golden.lengths))
        assert comp is not None and comp[0][-1] == rc.chrom_id.shape[0]
        loaded_ctx.reads_upload_compact(*comp)
        got = loaded_ctx.pileup(c, s, e)
        assert np.array_equal(got[0], base_pile[0]) and np.array_equal(got[1], base_pile[1])
        assert loaded_ctx.count_reads(c, s, e) == base_count
        assert lib.compact_read_columns(rc.chrom_id[::-1], rc.ref_start, rc.ref_end, rc.clip5(), flags, len(golden.lengths)) is None
        with pytest.raises(lib.TobiasB200Error):
            loaded_ctx.genome_upload_packed_runs(0, pack_ascii(golden.genome[0])[0], [5], [golden.lengths[0] + 1], golden.lengths[0])
    finally:
        loaded_ctx.genome_load(golden.genome)
        loaded_ctx.reads_upload_columns(rc)


def _set_model(ctx, g, mode):
    for si, st in enumerate(("forward", "reverse")):
        if mode == "DWM":
            ctx.set_bias_model(si, True, 25, *[g["DWM_%s_%s" % (st, k)] for k in ("bias_pwm_log", "bg_pwm_log", "bias_dwm_log", "bg_dwm_log")])
        else:
            ctx.set_bias_model(si, False, 25, g["PWM_%s_pssm" % st])


def _oracle_matrix(g, mode, st):
    m = restate.Matrix(25, mode)
    if mode == "DWM":
        for k in ("bias_pwm_log", "bg_pwm_log", "bias_dwm_log", "bg_dwm_log"):
            setattr(m, k, g["DWM_%s_%s" % (st, k)])
    else:
        m.pssm = g["PWM_%s_pssm" % st]
    return m


@pytest.mark.parametrize("mode", ["DWM", "PWM"])
def test_score_sequence(loaded_ctx, golden, mode):
    g = golden.atac
    _set_model(loaded_ctx, g, mode)
    for (c, a, b) in ((0, 700, 2300), (1, 12000, 13100), (0, 0, 60)):       # spans N blocks, soft-masked bases, a contig start
        codes = restate.encode(golden.genome[c][a:b])
        for si, st in enumerate(("forward", "reverse")):
            m = _oracle_matrix(g, mode, st)
            ref = m.score_sequence(codes) if si == 0 else m.score_sequence(restate.revcomp(codes))[::-1]
            got = loaded_ctx.score_sequence(si, c, a, b)
            assert np.array_equal(np.isnan(ref), np.isnan(got))
            v = ~np.isnan(ref)
            if mode == "PWM":
                assert np.array_equal(ref[v], got[v])
            else:
                assert np.all(np.abs(ref[v] - got[v]) <= 1e-9 * np.maximum(1.0, np.abs(ref[v])))


@pytest.mark.parametrize("L", [5, 21, 31])
def test_score_sequence_other_model_lengths(loaded_ctx, golden, L):
    """--k_flank other than the default 12: the product-form kernel picks another row grouping (run-time loop bounds instead of the
    unrolled L = 25 form; for L = 31 every lane owns a column).  Models built by the oracle's prepare_mat from random k-mer counts."""
    g = golden.atac
    rng = np.random.default_rng(L)
    models = []
    for si in range(2):
        m = restate.Matrix(L, "DWM")
        for _ in range(400):
            m.add_sequence(rng.integers(0, 4, size=L), float(rng.integers(1, 11)))
            m.add_background(rng.integers(0, 4, size=L), float(rng.integers(1, 11)))
        m.prepare_mat()
        models.append(m)
        loaded_ctx.set_bias_model(si, True, L, m.bias_pwm_log, m.bg_pwm_log, m.bias_dwm_log, m.bg_dwm_log)
        assert loaded_ctx.bias_model_info(si) == (1, L, 1)
    for (c, a, b) in ((0, 700, 1500), (1, 12000, 12300), (0, 0, 70)):
        codes = restate.encode(golden.genome[c][a:b])
        for si, m in enumerate(models):
            ref = m.score_sequence(codes) if si == 0 else m.score_sequence(restate.revcomp(codes))[::-1]
            got = loaded_ctx.score_sequence(si, c, a, b)
            assert np.array_equal(np.isnan(ref), np.isnan(got))
            v = ~np.isnan(ref)
            assert np.all(np.abs(ref[v] - got[v]) <= 1e-9 * np.maximum(1.0, np.abs(ref[v])))
    _set_model(loaded_ctx, g, "DWM")


@pytest.mark.parametrize("spread", [4.0, 12.0])
def test_score_sequence_table_range(loaded_ctx, golden, spread):
    """Product-form DWM scoring multiplies powers of two of the table rows.  spread 4: column products up to ~2^200, still
    the product form; spread 12: the bound on a column product passes the 2^480 the kernel allows, tb_set_bias_model must route the model to the sum form
    (sequences.pyx:263-342 evaluated with exp2 / log2 per column).  Both agree with the oracle."""
    g = golden.atac
    rng = np.random.default_rng(int(spread))
    tabs = {}
    for st in ("forward", "reverse"):
        for k in ("bias_pwm_log", "bg_pwm_log", "bias_dwm_log", "bg_dwm_log"):
            base = np.array(g["DWM_%s_%s" % (st, k)], dtype=np.float64)
            tabs[(st, k)] = base + rng.uniform(-spread, spread, size=base.shape)
    for si, st in enumerate(("forward", "reverse")):
        loaded_ctx.set_bias_model(si, True, 25, tabs[(st, "bias_pwm_log")], tabs[(st, "bg_pwm_log")], tabs[(st, "bias_dwm_log")],
                                  tabs[(st, "bg_dwm_log")])
        assert loaded_ctx.bias_model_info(si) == (1, 25, 1 if spread < 10 else 0)
    c, a, b = 0, 700, 1500
    codes = restate.encode(golden.genome[c][a:b])
    for si, st in enumerate(("forward", "reverse")):
        m = restate.Matrix(25, "DWM")
        for k in ("bias_pwm_log", "bg_pwm_log", "bias_dwm_log", "bg_dwm_log"):
            setattr(m, k, tabs[(st, k)])
        ref = m.score_sequence(codes) if si == 0 else m.score_sequence(restate.revcomp(codes))[::-1]
        got = loaded_ctx.score_sequence(si, c, a, b)
        assert np.array_equal(np.isnan(ref), np.isnan(got))
        v = np.isfinite(ref)
        assert v.sum() > 500
        assert np.all(np.abs(ref[v] - got[v]) <= 1e-9 * np.maximum(1.0, np.abs(ref[v])))
    _set_model(loaded_ctx, g, "DWM")          # leave the session context as the other tests expect it


@pytest.mark.parametrize("mode,lm", [("PWM", 1), ("DWM", 1), ("DWM", 0)])
def test_bias_correction_tracks(loaded_ctx, golden, mode, lm):
    """lm = 1: MINPACK-exact window fits (PWM default, --lm-exact); lm = 0: the one-pass form, the DWM default and the benchmarked one."""
    g = golden.atac
    _set_model(loaded_ctx, g, mode)
    c, s, e = _cols(g["outreg"])
    loaded_ctx.correct_run(c, s, e, correction_factor=float(g["correction_factor"]), lm_exact_order=lm)
    got = loaded_ctx.correct_download(split_strands=True, as_f64=True)
    for t in TRACKS:
        for si, st in enumerate(("forward", "reverse")):
            ref = g["%s_track_%s_%s" % (mode, t, st)]
            x = got[t][si]
            assert x.shape == ref.shape
            if t == "uncorrected" or mode == "PWM":
                assert np.array_equal(x, ref), (mode, t, st, np.abs(x - ref).max())
            elif t == "bias":
                assert np.all(np.abs(x - ref) <= 1e-9 * np.maximum(1.0, np.abs(ref)))
            else:
                d = np.abs(x - ref)
                ok = d <= 1e-5 * np.abs(ref) + 2e-6
                assert ok.mean() >= 0.999, (mode, t, st, ok.mean(), d.max())
                assert d.max() <= 2e-5, (mode, t, st, d.max())
    pp = got["prepost"]
    rtol = 1e-12 if mode == "PWM" else 1e-6
    assert np.allclose(pp[:, :, :4, :], g[mode + "_prepost"][:, :, :4, :], rtol=rtol, atol=1e-9)
    # what the bigWig writer stores: float32, strands summed
    both = loaded_ctx.correct_download(split_strands=False, as_f64=False, prepost=False)
    for t in TRACKS:
        ref32 = (g["%s_track_%s_forward" % (mode, t)] + g["%s_track_%s_reverse" % (mode, t)]).astype(np.float32)
        assert both[t].dtype == np.float32
        if mode == "PWM" or t == "uncorrected":
            assert np.array_equal(both[t], ref32)
        else:
            assert np.all(np.abs(both[t].astype(np.float64) - ref32) <= 1e-5 * np.abs(ref32) + 2e-5)


@pytest.mark.parametrize("lm_mode", [0, 5, 2, 3])
def test_bias_correction_other_fit_kernels(loaded_ctx, golden, lm_mode):
    """lm_exact_order = 0: one-pass (inner product) outer iterations, the DWM default; 5: the moment form (opt-in); 2 / 3:
    first-generation warp-per-window kernels (warp-tree sums / MINPACK-order sums), compiled into test / comparison builds only
    (-DTB_FIRST_GEN_FITS; the product library refuses them): same tracks within the DWM-mode tolerance even on the bit-exact PWM bias
    input (bit-identical for mode 3)."""
    g = golden.atac
    if lm_mode in (2, 3) and not loaded_ctx.device_info()["is_emulator"]:
        with pytest.raises(lib.TobiasB200Error, match="not part of this build"):
            loaded_ctx.correct_run(*_cols(g["outreg"]), correction_factor=1.0, lm_exact_order=lm_mode)
        return
    _set_model(loaded_ctx, g, "PWM")
    c, s, e = _cols(g["outreg"])
    loaded_ctx.correct_run(c, s, e, correction_factor=float(g["correction_factor"]), lm_exact_order=lm_mode)
    got = loaded_ctx.correct_download(split_strands=True, as_f64=True, prepost=False)
    for t in ("expected", "corrected"):
        for si, st in enumerate(("forward", "reverse")):
            ref = g["PWM_track_%s_%s" % (t, st)]
            d = np.abs(got[t][si] - ref)
            assert (d <= 1e-5 * np.abs(ref) + 2e-6).mean() >= 0.999 and d.max() <= 2e-5
            if lm_mode == 3:
                assert np.array_equal(got[t][si], ref)


def test_persistent_fit_kernel_equals_the_round_form(loaded_ctx, golden):
    """Forward-difference one-pass fits (lm_exact_order = 4, the first one-pass form): the persistent kernel (tile of the window data in
    shared memory, state in registers, lanes share every pass; option k3b_persistent) runs the same arithmetic in the same order as the
    round kernels -- fit records and tracks bit-identical, also when fits are parked after 1 or 3 passes and finished by the round
    kernels, and with tiles of 64 slots.  (Measured slower than the rounds on the B200, hence off by default.)"""
    g = golden.atac
    _set_model(loaded_ctx, g, "DWM")
    c, s, e = _cols(g["outreg"])
    cf = float(g["correction_factor"])
    loaded_ctx.correct_run(c, s, e, correction_factor=cf, lm_exact_order=4)
    base_trace = loaded_ctx.correct_fit_trace()
    base = loaded_ctx.correct_download(split_strands=True, as_f64=True, prepost=False)
    loaded_ctx.set_option("k3b_persistent", 1)
    try:
        for cap, tile in ((-1, -1), (1, -1), (3, 64)):
            loaded_ctx.set_option("k3b_cap", cap)
            loaded_ctx.set_option("k3b_tile", tile)
            loaded_ctx.correct_run(c, s, e, correction_factor=cf, lm_exact_order=4)
            assert np.array_equal(loaded_ctx.correct_fit_trace(), base_trace), (cap, tile)
            got = loaded_ctx.correct_download(split_strands=True, as_f64=True, prepost=False)
            for t in TRACKS:
                for si in range(2):
                    assert np.array_equal(got[t][si], base[t][si]), (cap, tile, t, si)
    finally:
        for k in ("k3b_cap", "k3b_tile", "k3b_persistent"):
            loaded_ctx.set_option(k, -1)
    assert (base_trace[:, :, 0] == 0).sum() > 500 and (base_trace[:, :, 5] > 25).sum() > 0      # fits long enough to be parked exist


def test_fit_forms_agree(loaded_ctx, golden):
    """The three arithmetic forms of the window fits -- MINPACK-exact (1), one pass with forward-difference inner products (0, the DWM
    default), one pass with the moments of the active set (5, opt-in) -- take the same fit / fallback / empty decision on every window of
    the golden data and agree on the fitted parameters to 1e-5."""
    g = golden.atac
    _set_model(loaded_ctx, g, "DWM")
    c, s, e = _cols(g["outreg"])
    tr = {}
    for lm in (1, 0, 5):
        loaded_ctx.correct_run(c, s, e, correction_factor=1.0, lm_exact_order=lm)
        tr[lm] = loaded_ctx.correct_fit_trace()
    for lm in (0, 5):
        assert np.array_equal(tr[lm][:, :, 0], tr[1][:, :, 0]), lm
        fitted = tr[1][:, :, 0] == 0
        for k in (1, 2, 3):
            d = np.abs(tr[lm][:, :, k] - tr[1][:, :, k])[fitted]
            assert d.max() <= 1e-5 * max(1.0, np.abs(tr[1][:, :, k][fitted]).max()), (lm, k, d.max())


def test_bias_correction_edge_regions(loaded_ctx, golden):
    """Regions the golden set does not hold, against the oracle region by region (PWM mode: every track bit-identical): a region whose
    extension starts at position 0 of a contig and lies in an N block without reads, one whose extension ends at the contig end, a
    one-base region, a region of another contig; and an empty region list."""
    g = golden.atac
    _set_model(loaded_ctx, g, "PWM")
    bias = restate.AtacBias(25, "PWM")
    for st in ("forward", "reverse"):
        bias.bias[st].pssm = g["PWM_%s_pssm" % st]
    bias.correction_factor = 0.75
    regs = [(0, 162, 700), (0, 5000, 5001), (0, 20000, 20350), (0, 41200, 41838), (1, 15000, 15900), (2, 400, 1400)]
    c, s, e = _cols(np.array(regs))
    loaded_ctx.correct_run(c, s, e, correction_factor=0.75, lm_exact_order=True)
    got = loaded_ctx.correct_download(split_strands=True, as_f64=True)
    off = 0
    pre_sum = np.zeros((2, 5, 25))
    for (ci, a, b) in regs:
        out, pre, post = restate.bias_correction_region(golden.reads, golden.genome, ci, a, b, bias)
        n = b - a
        for t in TRACKS:
            for si, st in enumerate(("forward", "reverse")):
                assert np.array_equal(got[t][si][off:off + n], out[t][st]), (ci, a, b, t, st)
        for si, st in enumerate(("forward", "reverse")):
            pre_sum[si] += pre[st].counts
        off += n
    assert off == got["corrected"][0].shape[0]
    assert np.allclose(got["prepost"][:, 0, :4, :], pre_sum[:, :4, :], rtol=1e-12, atol=1e-9)
    # no regions at all: a valid, empty result
    z = np.zeros(0, dtype=np.int64)
    loaded_ctx.correct_run(z.astype(np.int32), z, z, correction_factor=1.0)
    empty = loaded_ctx.correct_download(split_strands=False, as_f64=False)
    assert all(empty[t].shape == (0,) for t in TRACKS)
    assert loaded_ctx.count_reads(z.astype(np.int32), z, z) == 0


def test_fit_trace_matches_scipy_exactly_in_pwm_mode(loaded_ctx, golden):
    g = golden.atac
    _set_model(loaded_ctx, g, "PWM")
    c, s, e = _cols(g["outreg"])
    loaded_ctx.correct_run(c[:3], s[:3], e[:3], correction_factor=1.0, lm_exact_order=True)
    trace = loaded_ctx.correct_fit_trace()
    bias = restate.AtacBias(25, "PWM")
    for st in ("forward", "reverse"):
        bias.bias[st].pssm = g["PWM_%s_pssm" % st]
    ref = {"forward": [], "reverse": []}
    for ci, a, b in zip(c[:3], s[:3], e[:3]):
        tr = {}
        restate.bias_correction_region(golden.reads, golden.genome, int(ci), int(a), int(b), bias, trace=tr)
        for st in ref:
            ref[st] += tr[st]
    for si, st in enumerate(("forward", "reverse")):
        r = np.array(ref[st])
        fitted = r[:, 4] == 0
        assert np.array_equal(trace[si, :, 0], r[:, 4])                       # fit / fallback / empty decisions
        assert np.array_equal(trace[si, fitted, 1], r[fitted, 0]) and np.array_equal(trace[si, fitted, 2], r[fitted, 1])
        assert np.array_equal(trace[si, fitted, 4], r[fitted, 2]) and np.array_equal(trace[si, fitted, 5], r[fitted, 3])


# ------------------------------------------------------------------------------------------------ K4
def test_rolling_bit_exact(ctx, golden):
    s = golden.signals
    b = s["roll_in"]
    for w in (1, 7, 100):
        assert np.array_equal(ctx.rolling(b, w, "sum"), s["roll_sum_%d" % w], equal_nan=True)
        assert np.array_equal(ctx.rolling(b, w, "mean"), s["roll_mean_%d" % w], equal_nan=True)
    long = np.random.default_rng(0).normal(size=5000)
    assert np.array_equal(ctx.rolling(long, 100, "sum"), restate.fast_rolling_math(long, 100, "sum"), equal_nan=True)


def _fp_close(got, ref):
    assert np.array_equal(got == 0, ref == 0)
    nz = ref != 0
    assert np.all(np.abs(got[nz] - ref[nz]) <= 1e-5 * np.abs(ref[nz]))


def test_footprint_against_golden(ctx, golden):
    s = golden.signals
    a = s["fp_in"].astype(np.float32)
    for tag in ("default", "flank100", "narrow"):
        fmin, fmax, pmin, pmax = [int(x) for x in s["fp_%s_args" % tag]]
        for as_f64 in (True, False):
            p = ctx.score_params("footprint", flank=0, flank_min=fmin, flank_max=fmax, fp_min=pmin, fp_max=pmax, as_f64=as_f64)
            got = ctx.score_regions(a, [a.shape[0]], p)
            _fp_close(got.astype(np.float64), s["fp_" + tag])
    p = ctx.score_params("FOS", flank=0, as_f64=True)
    got = ctx.score_regions(a, [a.shape[0]], p)
    ref = s["fos_default"]
    assert np.array_equal(got == 10000, ref == 10000)
    assert np.allclose(got, ref, rtol=1e-5)


def test_footprint_many_regions_and_tiles(ctx):
    rng = np.random.default_rng(5)
    lens = [61, 140, 2100, 999, 141]                       # shorter than one scan, exactly one position, multi-tile
    sig = [np.where(rng.random(n) < 0.5, 0, rng.standard_t(3, size=n)).astype(np.float32) for n in lens]
    flat = np.concatenate(sig)
    flat[5] = np.nan
    p = ctx.score_params("footprint", flank=30, as_f64=True)
    got = ctx.score_regions(flat, lens, p)
    off = 0
    for i, n in enumerate(lens):
        x = flat[sum(lens[:i]):sum(lens[:i]) + n]
        ref = restate.calculate_scores_region(x, 30, "footprint")
        _fp_close(got[off:off + n - 60], ref)
        off += n - 60
    assert off == got.shape[0]


@pytest.mark.parametrize("form", [0, 32])
def test_footprint_forms_agree(ctx, golden, form):
    """Option k4_form: the first form of the footprint scan (one start per thread, 0) and the single-pass variant of the register-tiled
    form (32) against the golden data and the default (register tile of 4 starts x 16 widths, two passes) -- same zero pattern, scores
    within rtol 1e-5 (all round fp64 prefix differences to float, in different places).  Forced block sizes move the tile borders."""
    s = golden.signals
    a = s["fp_in"].astype(np.float32)
    rng = np.random.default_rng(8)
    lens = [760] * 4 + [61, 140, 2100, 91, 92, 5000]
    sig = [np.where(rng.random(n) < 0.5, 0, rng.standard_t(3, size=n)).astype(np.float32) for n in lens]
    try:
        ctx.set_option("k4_form", form)
        for tag in ("default", "flank100", "narrow"):
            fmin, fmax, pmin, pmax = [int(x) for x in s["fp_%s_args" % tag]]
            p = ctx.score_params("footprint", flank=0, flank_min=fmin, flank_max=fmax, fp_min=pmin, fp_max=pmax, as_f64=True)
            _fp_close(ctx.score_regions(a, [a.shape[0]], p), s["fp_" + tag])
        p = ctx.score_params("footprint", flank=30, as_f64=True)
        other = ctx.score_regions(np.concatenate(sig), lens, p)
    finally:
        ctx.set_option("k4_form", -1)
    outs = [ctx.score_regions(np.concatenate(sig), lens, p)]
    try:
        for nthr in (64, 160):
            ctx.set_option("k4_threads", nthr)
            outs.append(ctx.score_regions(np.concatenate(sig), lens, p))
    finally:
        ctx.set_option("k4_threads", -1)
    for o in outs[1:]:                                     # other tile borders: other reference points of the fp32 window sums
        assert np.array_equal(o == 0, outs[0] == 0) and np.all(np.abs(o - outs[0]) <= 2e-6 * np.abs(outs[0]))
    assert np.array_equal(other == 0, outs[0] == 0) and np.all(np.abs(other - outs[0]) <= 1e-5 * np.abs(outs[0]))


def test_footprint_width_ranges_and_tile_sizes(ctx):
    """more than 32 footprint widths (second batch: per-width sliding maximum), a single width, regions of very different
    lengths (the host picks the block size from them) -- against the oracle's restatement of tobias_footprint_array."""
    rng = np.random.default_rng(11)
    for lens, args in (([900, 180, 2500], (5, 12, 10, 60)), ([760] * 5, (10, 30, 20, 50)), ([300, 4000], (3, 3, 7, 7))):
        sig = [np.where(rng.random(n) < 0.6, 0, rng.normal(0, 1.0, size=n)).round(4).astype(np.float32) for n in lens]
        p = ctx.score_params("footprint", flank=0, flank_min=args[0], flank_max=args[1], fp_min=args[2], fp_max=args[3], as_f64=True)
        got = ctx.score_regions(np.concatenate(sig), lens, p)
        off = 0
        for x in sig:
            ref = restate.tobias_footprint_array(x.astype(np.float64), *args)
            _fp_close(got[off:off + x.shape[0]], ref)
            off += x.shape[0]


def test_bindetect_site_and_background_scores(ctx):
    """The per-base step of BINDetect's scan_and_score (bindetect_functions.py:294-331): footprint scores at the middle of every
    site and at the region-length-seeded random background positions, gathered on the device from the track of the last
    score_run -- against indexing the oracle's track the way the reference indexes the bigWig values."""
    import random
    from tobias_b200 import bindetect_functions as bf
    rng = np.random.default_rng(21)
    lens = [460, 1300, 260, 5000]                          # extended by flank 30 on both sides
    starts = [10_000, 20_000, 30_000, 40_000]              # genomic start of every trimmed region
    sig = [np.where(rng.random(n) < 0.5, 0, rng.standard_t(3, size=n)).astype(np.float32) for n in lens]
    p = ctx.score_params("footprint", flank=30)
    ctx.signal_upload(np.concatenate(sig), lens)
    ctx.score_run(p)
    track = [restate.calculate_scores_region(x, 30, "footprint") for x in sig]
    sites = []
    for r, n in enumerate(lens):
        for _ in range(40):
            width = int(rng.integers(6, 25))
            a = starts[r] + int(rng.integers(0, n - 60 - width))
            sites.append((r, a, a + width))
    got_sites, got_bg, bg_off = bf.gather_scores(ctx, [n - 60 for n in lens], starts, sites)
    full = ctx.score_download()                            # float32 track the bigWig would hold
    off = np.concatenate([[0], np.cumsum([n - 60 for n in lens])])
    for i, (r, a, b) in enumerate(sites):
        pos = a - starts[r] + int((b - a) / 2.0)           # :323-324
        assert got_sites[i] == float(full[off[r] + pos])
        assert abs(got_sites[i] - track[r][pos]) <= 1e-5 * max(1.0, abs(track[r][pos]))
    k = 0
    for r, n in enumerate(lens):
        reglen = n - 60
        random.seed(reglen)                                # :295-296
        want = random.sample(range(reglen), max(1, int(reglen / 200)))
        assert bg_off[r] == want
        for pos in want:
            assert got_bg[k] == float(full[off[r] + pos])
            k += 1
    assert k == got_bg.shape[0]
    with pytest.raises(Exception):
        ctx.score_gather([int(off[-1])])                   # one past the end of the track


def test_calculate_scores_against_golden(ctx, golden):
    sc, g = golden.score, golden.atac
    sig = sc["signal"]
    regs = [tuple(int(x) for x in r) for r in g["outreg"]]
    opts = {
        "footprint": dict(kind="footprint"), "footprint_smooth": dict(kind="footprint", smooth=5),
        "footprint_limits": dict(kind="footprint", min_limit=-0.2, max_limit=0.6),
        "sum_abs": dict(kind="sum", absolute=True, window=100), "mean": dict(kind="mean", window=21), "none": dict(kind="none"),
    }
    for tag, opt in opts.items():
        flank = int(sc["flank_" + tag])
        ext, lens = [], []
        for (c, s, e) in regs:
            x = np.zeros(e - s + 2 * flank, dtype=np.float32)
            o2 = 0
            for (c2, s2, e2) in regs:
                if c2 == c:
                    a, b = max(s - flank, s2), min(e + flank, e2)
                    if a < b:
                        x[a - (s - flank):b - (s - flank)] = sig[o2 + a - s2:o2 + b - s2]
                o2 += e2 - s2
            ext.append(x)
            lens.append(x.shape[0])
        got = ctx.score_regions(np.concatenate(ext), lens, ctx.score_params(flank=flank, as_f64=True, **opt))
        ref = sc["out_" + tag]
        assert got.shape == ref.shape
        if opt["kind"] == "footprint":
            assert np.array_equal(np.isnan(got), np.isnan(ref))
            v = ~np.isnan(ref)
            assert np.all(np.abs(got[v] - ref[v]) <= 1e-5 * np.abs(ref[v]) + 1e-12), tag
        else:
            assert np.array_equal(got, ref, equal_nan=True), tag


def test_chain_corrected_to_footprints(loaded_ctx, golden):
    """ATACorrect -> ScoreBigwig chained on the device equals scoring the float32 corrected track."""
    g, sc = golden.atac, golden.score
    _set_model(loaded_ctx, g, "DWM")
    c, s, e = _cols(g["outreg"])
    loaded_ctx.correct_run(c, s, e, correction_factor=float(g["correction_factor"]))
    both = loaded_ctx.correct_download(split_strands=False, as_f64=False, tracks=("corrected",), prepost=False)["corrected"]
    loaded_ctx.signal_from_corrected(30, e - s)
    loaded_ctx.score_run(loaded_ctx.score_params("footprint", flank=30, as_f64=True))
    got = loaded_ctx.score_download()
    ext, lens, off = [], [], 0
    for a, b in zip(s, e):
        x = np.zeros(b - a + 60, dtype=np.float32)
        x[30:-30] = both[off:off + b - a]
        off += b - a
        ext.append(x)
        lens.append(x.shape[0])
    ref = loaded_ctx.score_regions(np.concatenate(ext), lens, loaded_ctx.score_params("footprint", flank=30, as_f64=True))
    assert np.array_equal(got, ref)
    # asynchronous download: copies enqueued before the footprint kernels, complete after the wait
    loaded_ctx.correct_run(c, s, e, correction_factor=float(g["correction_factor"]))
    later = loaded_ctx.correct_download(split_strands=False, as_f64=False, tracks=("corrected", "expected"), prepost=True, wait=False)
    loaded_ctx.signal_from_corrected(30, e - s)
    loaded_ctx.score_run(loaded_ctx.score_params("footprint", flank=30, as_f64=True))
    loaded_ctx.correct_download_wait()
    assert np.array_equal(later["corrected"], both) and np.array_equal(loaded_ctx.score_download(), got)


def test_streaming_sinks(loaded_ctx, golden):
    """tb_correct_set_sink / tb_score_set_sink: the run hands its tracks (chunks of output regions, last stage launched per chunk) and the
    footprint scan its scores to the host while the next chunk computes -- the same arrays as the downloads after the run, for every
    chunk count, and a sink serves one run only."""
    g = golden.atac
    _set_model(loaded_ctx, g, "DWM")
    c, s, e = _cols(g["outreg"])
    n_out = int((e - s).sum())
    cf = float(g["correction_factor"])
    p30 = loaded_ctx.score_params("footprint", flank=30)
    loaded_ctx.correct_run(c, s, e, correction_factor=cf, lm_exact_order=0)
    base = loaded_ctx.correct_download(split_strands=False, as_f64=False, prepost=True)
    loaded_ctx.signal_from_corrected(30, e - s)
    loaded_ctx.score_run(p30)
    fp = loaded_ctx.score_download()
    try:
        for chunks in (1, 2, 6):
            loaded_ctx.set_option("k3c_chunks", chunks)
            loaded_ctx.set_option("k4_chunks", chunks)
            bufs = {t: loaded_ctx.pinned_empty((n_out + 5,), np.float32) for t in ("uncorrected", "bias", "expected", "corrected")}
            for b in bufs.values():
                b[:] = -7.0
            views = loaded_ctx.correct_set_sink(bufs, n_out, as_f64=False, prepost=True)
            loaded_ctx.correct_run(c, s, e, correction_factor=cf, lm_exact_order=0)
            loaded_ctx.signal_from_corrected(30, e - s)
            out = loaded_ctx.pinned_empty((fp.shape[0] + 3,), np.float32)
            out[:] = -7.0
            loaded_ctx.score_set_sink(out)
            loaded_ctx.score_run(p30)
            loaded_ctx.correct_download_wait()
            loaded_ctx.score_sink_wait()
            for t in bufs:
                assert np.array_equal(views[t], base[t], equal_nan=True), (chunks, t)
                assert np.all(bufs[t][n_out:] == -7.0)
            assert np.allclose(views["prepost"], base["prepost"], rtol=1e-12, atol=1e-9)      # fp64 atomics: the order of the tiles differs
            assert np.array_equal(out[:fp.shape[0]], fp) and np.all(out[fp.shape[0]:] == -7.0)
    finally:
        loaded_ctx.set_option("k3c_chunks", -1)
        loaded_ctx.set_option("k4_chunks", -1)
    # the sinks are spent: the next runs leave the buffers alone
    for b in bufs.values():
        b[:] = -7.0
    out[:] = -7.0
    loaded_ctx.correct_run(c, s, e, correction_factor=cf, lm_exact_order=0)
    loaded_ctx.signal_from_corrected(30, e - s)
    loaded_ctx.score_run(p30)
    loaded_ctx.correct_download_wait()
    assert all(np.all(b == -7.0) for b in bufs.values()) and np.all(out == -7.0)
    assert np.array_equal(loaded_ctx.score_download(), fp)


def test_chain_reads_the_track_like_a_bigwig(loaded_ctx, golden):
    """tb_signal_from_corrected: the flank of an output region takes the values of an abutting neighbour (50-kb split pieces) and 0
    where the track holds nothing -- what ScoreBigwig reads from the corrected bigWig; tb_signal_from_corrected_regions scores regions
    of one's own choice (here: the merged region, as score_bigwig.py:136-170 would build it) over the same track."""
    g = golden.atac
    _set_model(loaded_ctx, g, "DWM")
    regs = np.array([(0, 20000, 20700), (0, 20700, 21300), (0, 21300, 21900), (0, 25000, 25600), (1, 9000, 9800)])
    c, s, e = _cols(regs)
    loaded_ctx.correct_run(c, s, e, correction_factor=float(g["correction_factor"]), lm_exact_order=0)
    both = loaded_ctx.correct_download(split_strands=False, as_f64=False, tracks=("corrected",), prepost=False)["corrected"]
    off = np.concatenate([[0], np.cumsum(e - s)])
    track = {}                                                   # (contig, position) -> value: the bigWig
    for i, (ci, a, b) in enumerate(regs):
        for k, p in enumerate(range(a, b)):
            track[(ci, p)] = both[off[i] + k]

    def read(ci, a, b):
        return np.array([track.get((ci, p), 0.0) for p in range(a, b)], dtype=np.float32)
    p30 = loaded_ctx.score_params("footprint", flank=30, as_f64=True)
    loaded_ctx.signal_from_corrected(30, e - s)
    loaded_ctx.score_run(p30)
    got = loaded_ctx.score_download()
    ext = [read(ci, a - 30, b + 30) for (ci, a, b) in regs]
    ref = loaded_ctx.score_regions(np.concatenate(ext), [x.shape[0] for x in ext], p30)
    assert np.array_equal(got, ref)
    assert np.any(ext[0][-30:] != 0) and np.all(ext[3][:30] == 0)          # a neighbour's values / nothing
    merged = np.array([(0, 20000, 21900), (0, 25000, 25600), (1, 9000, 9800)])
    mc, ms, me = _cols(merged)
    loaded_ctx.signal_from_corrected_regions(30, mc, ms, me)
    loaded_ctx.score_run(p30)
    got_m = loaded_ctx.score_download()
    ext_m = [read(ci, a - 30, b + 30) for (ci, a, b) in merged]
    ref_m = loaded_ctx.score_regions(np.concatenate(ext_m), [x.shape[0] for x in ext_m], p30)
    assert np.array_equal(got_m, ref_m) and got_m.shape[0] == int((me - ms).sum())
    with pytest.raises(lib.TobiasB200Error):
        loaded_ctx.signal_from_corrected_regions(30, [0], [10], [500])             # flank leaves the contig
