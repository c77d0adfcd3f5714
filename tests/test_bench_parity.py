"""Parity of the BENCHMARKED configuration against the unmodified reference (oracle/_ref), through the same code the bench
line's `parity` block runs (bench.cpu_reference(collect=True) + bench.compare_with_reference):

  * emu (CPU):  a tiny workload through the kernel emulator -- checks the comparison machinery itself (region / window
                alignment, fit records, degenerate-window accounting) and the default DWM path (lm_exact_order = 0).
  * gpu:        BASELINE.json configs[0] (one 50 Mb contig, 20 k peaks, 5 M records) built with synth_torch; bias model from
                ALL input regions (count tensors bit-exact against the reference's bias_estimation over the whole contig),
                >= 2000 sampled output regions compared track by track, DWM with lm_exact_order = 0 AND 1.

Tolerances: bench.TOL (stated there and in the bench line)."""
import os
import types

import pytest

import bench


def _tiny(name):
    return dict(lengths={"chr1": 160000, "chr2": 120000, "chrM": 16569}, n_peaks=16, n_records=40000, desc="tiny (parity machinery test)")


def _check(par, need_green):
    assert par["read_counts_exact"] and par["count_tensors_exact"] and par["log_tables_exact"], par
    fit = par["lm_windows"]
    assert fit["regions_with_call_count_mismatch"] == 0 and fit["stable_decision_flips"] == 0, fit
    assert fit["stable_mismatches"] <= fit["stable_mismatch_budget"], fit
    assert fit["fitted_ref"] + fit["fallback_ref"] + fit["empty"] == fit["windows"]
    assert par["tracks"]["uncorrected"]["max_abs"] == 0.0
    assert par["tracks"]["bias"]["ok"], par["tracks"]["bias"]
    assert fit["unstable_flips"] + fit["stable_mismatches"] == fit["mismatching_windows"]
    if need_green:
        assert par["green"], {k: v for k, v in par.items() if k != "tolerance"}


def test_parity_block_on_the_emulator(emu_lib, monkeypatch):
    from tobias_b200.pipeline import Engine
    monkeypatch.setattr(bench, "workload", _tiny)
    args = types.SimpleNamespace(workload="tiny")
    ref = bench.cpu_reference(args, steps=1, warmup=0, n_out=8, collect=True)
    assert ref["cpu_baseline"]["kind"] == "reference" and ref["cpu_baseline"]["value"] > 0
    eng = Engine(0, emu_lib)
    try:
        par = bench.compare_with_reference(eng, ref["sample"], ref["kept"], 0)
    finally:
        eng.close()
    assert par["regions_compared"] == 8 and par["bp_compared"] > 5000
    _check(par, need_green=False)
    for t in ("expected", "corrected", "footprint"):          # tiny sample: the 99.9 % rule is too coarse here, the 10x rule is not
        assert par["tracks"][t]["outside_x10_unexplained"] <= 2, par["tracks"][t]


@pytest.mark.gpu
def test_chr50mb_against_the_reference_both_fit_modes():
    """configs[0] at full size: the whole 50 Mb contig through counts -> model, >= 2000 output regions against oracle/_ref."""
    from tobias_b200.pipeline import Engine
    args = types.SimpleNamespace(workload="chr50mb")
    ref = bench.cpu_reference(args, steps=1, warmup=0, n_out=max(2048, (os.cpu_count() or 1) * 64), collect=True)
    eng = Engine(0)
    try:
        for lm in (0, 1):
            par = bench.compare_with_reference(eng, ref["sample"], ref["kept"], lm)
            assert par["regions_compared"] >= 2000
            print("chr50mb parity lm_exact_order=%d: %s" % (lm, {k: v for k, v in par.items() if k not in ("tolerance", "note")}))
            _check(par, need_green=True)
    finally:
        eng.close()
