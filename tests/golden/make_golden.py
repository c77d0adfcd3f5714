#!/usr/bin/env python
"""
Generates the golden fixtures of tests/golden/ by running the UNMODIFIED reference (compiled into oracle/_ref by
oracle/build_ref.py from /root/reference) on a tiny seeded synthetic dataset.  Run in the build container (needs
/root/reference or a prebuilt oracle/_ref); the resulting .npz files are committed and are what the CPU and GPU
test-suites compare against on machines where the reference tree does not exist.

    python tests/golden/make_golden.py

Fixtures
  golden_dataset.npz     the inputs: genome contigs (ASCII), read columns, peaks
  golden_signals.npz     fast_rolling_math / tobias_footprint_array / FOS_score in/out        (signals.pyx)
  golden_atacorrect.npz  count_reads, bias_estimation counts (DWM + PWM), prepare_mat tables, bias_correction tracks,
                         pre/post verification counts for DWM and PWM                       (atacorrect_functions.py)
  golden_scorebigwig.npz calculate_scores outputs for several option sets                     (score_bigwig.py)
"""
import os
import sys
import tempfile
import types
from copy import deepcopy

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tobias_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402

SEED = 11
LENGTHS = {"chrA": 42000, "chrB": 31000, "chrM": 5000}


def dataset():
    return synth.SynthDataset(SEED, LENGTHS, n_peaks=10, n_records=26000)


class _Q:
    def __init__(self):
        self.items = []

    def put(self, x):
        self.items.append(x)


def main():
    ref = ref_loader.load()
    R = ref.regions
    ds = dataset()
    names = list(ds.genome.keys())
    lens = [len(ds.genome[n]) for n in names]
    tmp = tempfile.mkdtemp(prefix="tb_golden_")
    paths = ds.write(tmp, bam=False)

    save = {}
    for n in names:
        save["genome_" + n] = ds.genome[n]
    for f in ("chrom_id", "ref_start", "ref_end", "left_soft", "right_soft", "query_length", "flag", "first_op", "last_op", "tlen"):
        save["reads_" + f] = getattr(ds.reads, f)
    save["contig_names"] = np.array(names)
    save["contig_lengths"] = np.array(lens, dtype=np.int64)
    save["peaks"] = np.array([(names.index(c), s, e) for c, s, e in ds.peaks], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "golden_dataset.npz"), **save)

    # ------------------------------------------------------------------ signals.pyx
    rng = np.random.default_rng(SEED)
    sig = {}
    a = np.round(rng.normal(0, 0.5, size=900), 5)
    a[rng.random(900) < 0.6] = 0
    a = a.astype(np.float32).astype(np.float64)          # bigWig values are float32
    sig["fp_in"] = a
    for tag, args in (("default", (10, 30, 20, 50)), ("flank100", (100, 100, 20, 50)), ("narrow", (3, 6, 4, 9))):
        sig["fp_" + tag] = ref.signals.tobias_footprint_array(a, *args)
        sig["fp_" + tag + "_args"] = np.array(args)
    sig["fos_default"] = ref.signals.FOS_score(a, 10, 30, 20, 50)
    b = rng.normal(0, 3, size=500)
    sig["roll_in"] = b
    for w in (1, 7, 100):
        sig["roll_sum_%d" % w] = ref.signals.fast_rolling_math(b, w, "sum")
        sig["roll_mean_%d" % w] = ref.signals.fast_rolling_math(b, w, "mean")
    np.savez_compressed(os.path.join(HERE, "golden_signals.npz"), **sig)

    # ------------------------------------------------------------------ ATACorrect worker functions
    peak_regions = R.RegionList().from_bed(paths["peaks"])
    peak_regions.merge()
    genome_regions = R.RegionList([R.OneRegion([n, 0, ln]) for n, ln in zip(names, lens) if n != "chrM"])
    nonpeak = deepcopy(genome_regions).subtract(peak_regions)
    nonpeak = nonpeak.apply_method(R.OneRegion.split_region, 9000)          # small pieces: reads straddle borders
    out_regions = deepcopy(peak_regions)
    out_regions.apply_method(R.OneRegion.extend_reg, 100 + 62)
    out_regions.merge()
    out_regions.apply_method(R.OneRegion.extend_reg, -62)
    g = {}
    g["nonpeak"] = np.array([(names.index(r.chrom), r.start, r.end) for r in nonpeak], dtype=np.int64)
    g["peakreg"] = np.array([(names.index(r.chrom), r.start, r.end) for r in peak_regions], dtype=np.int64)
    g["outreg"] = np.array([(names.index(r.chrom), r.start, r.end) for r in out_regions], dtype=np.int64)
    for mode in ("DWM", "PWM"):
        params = types.SimpleNamespace(bam=paths["reads_npz"], genome=paths["fasta"], k_flank=12, bg_shift=100,
                                       read_shift=[4, -5], score_mat=mode, verbosity=0, log_q=None, window=100,
                                       split_strands=True)
        if mode == "DWM":
            g["count_nonpeak"] = np.array(ref.atacorrect_functions.count_reads(deepcopy(nonpeak), params))
            g["count_peak"] = np.array(ref.atacorrect_functions.count_reads(deepcopy(peak_regions), params))
        bias = ref.atacorrect_functions.bias_estimation(deepcopy(nonpeak), params)
        g[mode + "_no_reads"] = np.array(bias.no_reads)
        for st in ("forward", "reverse"):
            m = bias.bias[st]
            g["%s_%s_counts" % (mode, st)] = m.counts.copy()
            g["%s_%s_bg_counts" % (mode, st)] = m.bg_counts.copy()
            g["%s_%s_no" % (mode, st)] = np.array([m.no_bias, m.no_bg], dtype=np.float64)
            m.prepare_mat()
            if mode == "DWM":
                for k in ("bias_pwm_log", "bg_pwm_log", "bias_dwm_log", "bg_dwm_log"):
                    g["%s_%s_%s" % (mode, st, k)] = getattr(m, k)
            g["%s_%s_pssm" % (mode, st)] = m.pssm
        bias.correction_factor = 1.2345
        q = _Q()
        params.qs = {"%s:%s" % (t, s): q for t in ("uncorrected", "bias", "expected", "corrected") for s in ("forward", "reverse")}
        pre, post = ref.atacorrect_functions.bias_correction(deepcopy(out_regions), params, bias)
        got = {(k, reg): arr for (k, reg, arr) in q.items}
        for t in ("uncorrected", "bias", "expected", "corrected"):
            for s in ("forward", "reverse"):
                g["%s_track_%s_%s" % (mode, t, s)] = np.concatenate(
                    [got[("%s:%s" % (t, s), (r.chrom, r.start, r.end))] for r in out_regions])
        g[mode + "_prepost"] = np.stack([np.stack([pre[s].counts, post[s].counts, post[s].neg_counts])
                                         for s in ("forward", "reverse")])
    g["correction_factor"] = np.array(1.2345)
    np.savez_compressed(os.path.join(HERE, "golden_atacorrect.npz"), **g)

    # ------------------------------------------------------------------ ScoreBigwig worker (signal = DWM corrected track, fp32)
    import pyBigWig  # the shim, backed by the repo's bigWig codec
    corrected = (g["DWM_track_corrected_forward"] + g["DWM_track_corrected_reverse"]).astype(np.float32)
    bw_path = os.path.join(tmp, "corrected.bw")
    bw = pyBigWig.open(bw_path, "w")
    bw.addHeader([(n, ln) for n, ln in zip(names, lens)])
    off = 0
    for r in out_regions:
        n = r.end - r.start
        vals = corrected[off:off + n]
        nz = np.flatnonzero(vals)
        if nz.size:
            bw.addEntries(r.chrom, (nz + r.start).tolist(), values=vals[nz].tolist(), span=1)
        off += n
    bw.close()
    s = {"signal": corrected}
    option_sets = {
        "footprint": dict(score="footprint", absolute=False, min_limit=None, max_limit=None, window=100, smooth=1),
        "footprint_smooth": dict(score="footprint", absolute=False, min_limit=None, max_limit=None, window=100, smooth=5),
        "footprint_limits": dict(score="footprint", absolute=False, min_limit=-0.2, max_limit=0.6, window=100, smooth=1),
        "sum_abs": dict(score="sum", absolute=True, min_limit=None, max_limit=None, window=100, smooth=1),
        "mean": dict(score="mean", absolute=False, min_limit=None, max_limit=None, window=21, smooth=1),
        "none": dict(score="none", absolute=False, min_limit=None, max_limit=None, window=100, smooth=1),
    }
    for tag, opt in option_sets.items():
        flank = {"sum": int(opt["window"] / 2.0), "footprint": 30}.get(opt["score"], 0)       # score_bigwig.py:154-159
        q = _Q()
        args = types.SimpleNamespace(signal=bw_path, verbosity=0, log_q=None, region_flank=flank, flank_min=10, flank_max=30,
                                     fp_min=20, fp_max=50, writer_qs={"scores": q}, **opt)
        ref.score_bigwig.calculate_scores(deepcopy(out_regions), args)
        got = {reg: arr for (_, reg, arr) in q.items}
        s["out_" + tag] = np.concatenate([got[(r.chrom, r.start, r.end)] for r in out_regions])
        s["flank_" + tag] = np.array(flank)
    np.savez_compressed(os.path.join(HERE, "golden_scorebigwig.npz"), **s)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
