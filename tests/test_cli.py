"""Drop-in check of the two command-line tools: the same arguments given to the reference's own drivers (compiled,
unmodified: oracle/_ref, run with its multiprocessing pools and bigwig writers over the pysam / pyBigWig shims) and to
tobias_b200's drivers must produce the same bigWig files.  CPU run: kernels execute in the emulator; `-m gpu`: on the B200."""
import argparse
import os
import sys

import numpy as np
import pytest

from tobias_b200 import synth
from tobias_b200.io.bigwig import BigWigReader


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli")
    ds = synth.SynthDataset(23, {"chr1": 60000, "chr2": 45000, "chrM": 5000}, n_peaks=12, n_records=40000)
    paths = ds.write(str(d), bam=True)
    return ds, paths, str(d)


def _args(parser_fn, argv):
    return parser_fn(argparse.ArgumentParser()).parse_args(argv)


def _bw_values(path, names_lengths):
    r = BigWigReader(path)
    out = {c: np.nan_to_num(r.values(c, 0, n)) for c, n in names_lengths if c in r.chroms}
    r.close()
    return out


def _reference_available():
    from oracle import build_ref
    return build_ref.is_built() or os.path.isdir("/root/reference/tobias")


@pytest.mark.parametrize("mode", ["PWM", "DWM"])
def test_atacorrect_and_scorebigwig_match_reference_cli(ctx, dataset, mode):
    if not _reference_available():
        pytest.skip("compiled reference not available")
    from oracle import ref_loader
    ds, paths, d = dataset
    ref_drv = ref_loader.load_driver()
    ref = ref_loader.load()
    common = ["--bam", paths["bam"], "--genome", paths["fasta"], "--peaks", paths["peaks"], "--score_mat", mode, "--verbosity", "0"]
    out_ref, out_mine = os.path.join(d, "ref_" + mode), os.path.join(d, "mine_" + mode)
    os.makedirs(out_ref, exist_ok=True)
    ref_drv.run_atacorrect(_args(ref.parsers.add_atacorrect_arguments, common + ["--outdir", out_ref, "--prefix", "x", "--cores", "2"]))
    from tobias_b200 import atacorrect, parsers, score_bigwig
    atacorrect.run_atacorrect(_args(parsers.add_atacorrect_arguments, common + ["--outdir", out_mine, "--prefix", "x"]))
    nl = [(n, len(ds.genome[n])) for n in ds.genome]
    for track in ("uncorrected", "bias", "expected", "corrected"):
        a = _bw_values(os.path.join(out_ref, "x_%s.bw" % track), nl)
        b = _bw_values(os.path.join(out_mine, "x_%s.bw" % track), nl)
        assert set(a) == set(b)
        for c in a:
            if mode == "PWM" or track == "uncorrected":
                assert np.array_equal(a[c], b[c]), (mode, track, c, np.abs(a[c] - b[c]).max())
            else:
                # DWM: rtol 1e-5 plus an absolute term relative to the track's scale (the tracks are multiplied by the
                # correction factor, ~1e3 on this tiny data set); see tests/test_kernels.py for why DWM mode is not bit-exact
                scale = max(float(np.abs(a[c]).max()), 1.0)
                diff = np.abs(a[c] - b[c])
                assert np.all(diff <= 1e-5 * np.abs(a[c]) + 2e-6 * scale), (mode, track, c, diff.max(), scale)
                assert np.mean(diff <= 1e-5 * np.abs(a[c]) + 1e-7 * scale) >= 0.999, (mode, track, c)
    assert os.path.exists(os.path.join(out_mine, "x_AtacBias.pickle"))
    pdf = open(os.path.join(out_mine, "x_atacorrect.pdf"), "rb").read()          # the report: 2 bias-motif pages + 2 pre / post pages
    assert pdf.startswith(b"%PDF-1.4") and pdf.count(b"/Type /Page ") == 4 and pdf.rstrip().endswith(b"%%EOF")
    assert b"Tn5 insertion bias of reads \\(forward\\)" in pdf and b"Nucleotide frequency" in pdf
    # ScoreBigwig on the reference's corrected track
    sig = os.path.join(out_ref, "x_corrected.bw")
    for extra in ([], ["--score", "sum", "--window", "50", "--absolute"], ["--smooth", "3", "--fp-max", "40"]):
        tag = "_".join(x.strip("-") for x in extra) or "default"
        o_ref, o_mine = os.path.join(d, "fp_ref_%s_%s.bw" % (mode, tag)), os.path.join(d, "fp_mine_%s_%s.bw" % (mode, tag))
        sb = ["--signal", sig, "--regions", paths["peaks"], "--verbosity", "0"] + extra
        ref.score_bigwig.run_scorebigwig(_args(ref.parsers.add_scorebigwig_arguments, sb + ["--output", o_ref, "--cores", "2"]))
        score_bigwig.run_scorebigwig(_args(parsers.add_scorebigwig_arguments, sb + ["--output", o_mine]))
        a, b = _bw_values(o_ref, nl), _bw_values(o_mine, nl)
        assert set(a) == set(b)
        for c in a:
            assert np.array_equal(a[c] == 0, b[c] == 0), (tag, c)
            assert np.all(np.abs(a[c] - b[c]) <= 1e-5 * np.abs(a[c]) + 1e-9), (tag, c, np.abs(a[c] - b[c]).max())


def test_cli_parser_matches_reference_flags():
    if not _reference_available():
        pytest.skip("compiled reference not available")
    from oracle import ref_loader
    from tobias_b200 import parsers
    ref = ref_loader.load()
    for name in ("add_atacorrect_arguments", "add_scorebigwig_arguments"):
        ref_p = getattr(ref.parsers, name)(argparse.ArgumentParser())
        my_p = getattr(parsers, name)(argparse.ArgumentParser())
        ref_opts = {o.dest: (tuple(o.option_strings), o.default, o.nargs, tuple(o.choices) if o.choices else None) for o in ref_p._actions}
        my_opts = {o.dest: (tuple(o.option_strings), o.default, o.nargs, tuple(o.choices) if o.choices else None) for o in my_p._actions}
        for dest, spec in ref_opts.items():
            assert my_opts.get(dest) == spec, (name, dest, spec, my_opts.get(dest))
        assert set(my_opts) - set(ref_opts) <= {"device", "lm_exact"}


def test_worker_function_mirrors(ctx, dataset):
    """count_reads / bias_estimation / bias_correction with the reference's signatures."""
    import types
    from copy import deepcopy
    from oracle import restate
    from tobias_b200 import atacorrect_functions as af
    from tobias_b200.regions import OneRegion, RegionList
    ds, paths, d = dataset
    names = list(ds.genome)
    params = types.SimpleNamespace(bam=paths["reads_npz"], genome=paths["fasta"], k_flank=12, bg_shift=100, read_shift=[4, -5],
                                   score_mat="PWM", window=100, split_strands=False, verbosity=0, log_q=None)
    peaks = RegionList([OneRegion(list(p)) for p in ds.peaks])
    peaks.merge()
    nonpeak = RegionList([OneRegion([n, 0, len(ds.genome[n])]) for n in names if n != "chrM"]).subtract(deepcopy(peaks))
    regs = [(names.index(r.chrom), r.start, r.end) for r in nonpeak]
    assert af.count_reads(nonpeak, params) == restate.count_reads(ds.reads, regs)
    bias = af.bias_estimation(nonpeak, params)
    ref_bias = restate.bias_estimation(ds.reads, [ds.genome[n] for n in names], regs, [len(ds.genome[n]) for n in names], score_mat="PWM")
    assert bias.no_reads == ref_bias.no_reads
    for st in ("forward", "reverse"):
        assert np.array_equal(bias.bias[st].counts, ref_bias.bias[st].counts)
        bias.bias[st].prepare_mat()
        ref_bias.bias[st].prepare_mat()
    bias.correction_factor = ref_bias.correction_factor = 2.5

    class Q:
        def __init__(self):
            self.items = []

        def put(self, x):
            self.items.append(x)
    q = Q()
    params.qs = {"corrected:both": q, "bias:both": q}
    out = RegionList(deepcopy(peaks)[:3])
    for r in out:
        r.extend_reg(100)
    pre, post = af.bias_correction(out, params, bias)
    assert len(q.items) == 6 and set(pre) == {"forward", "reverse"}
    for (key, reg, sig) in q.items:
        o, _, _ = restate.bias_correction_region(ds.reads, [ds.genome[n] for n in names], names.index(reg[0]), reg[1], reg[2], ref_bias)
        t = key.split(":")[0]
        assert np.array_equal(sig, o[t]["forward"] + o[t]["reverse"]), key
    af.release_engines()
