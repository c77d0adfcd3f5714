"""<prefix>_AtacBias.pickle compatibility with the reference (tobias/tools/atacorrect_functions.py:62-76; `--bias-pkl`,
tobias/tools/atacorrect.py:316-332): pickles written here load in the reference, pickles written by the reference load here,
attribute for attribute -- checked against the compiled reference (oracle/_ref)."""
import os
import pickle
import sys

import numpy as np
import pytest

from tobias_b200.pipeline import AtacBias

ATTRS_DWM = ("counts", "bg_counts", "neg_counts", "pssm", "bias_pwm", "bg_pwm", "bias_dwm", "bg_dwm", "bias_pwm_log", "bg_pwm_log",
             "bias_dwm_log", "bg_dwm_log")
ATTRS_PWM = ("counts", "bg_counts", "neg_counts", "pssm", "bias_pwm", "bg_pwm")


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_loader
    return ref_loader.load()


def _fill(obj, stype, rng):
    for st in ("forward", "reverse"):
        m = obj.bias[st]
        shape = m.counts.shape
        m.counts = rng.integers(0, 50, size=shape).astype(np.float64)
        m.bg_counts = rng.integers(0, 50, size=shape).astype(np.float64)
        if stype == "DWM":
            m.counts[4] = m.counts[:, 4] = 0
            m.bg_counts[4] = m.bg_counts[:, 4] = 0
        m.prepare_mat()
    obj.no_reads = 12345
    obj.correction_factor = 0.8125


def _same(a, b, stype):
    assert a.stype == b.stype == stype and a.no_reads == b.no_reads and a.correction_factor == b.correction_factor
    for st in ("forward", "reverse", "both"):
        ma, mb = a.bias[st], b.bias[st]
        assert ma.length == mb.length and ma.no_bias == mb.no_bias and ma.no_bg == mb.no_bg
        for k in (ATTRS_DWM if stype == "DWM" else ATTRS_PWM):
            if hasattr(ma, k) or hasattr(mb, k):
                assert np.array_equal(np.asarray(getattr(ma, k)), np.asarray(getattr(mb, k)), equal_nan=True), (st, k)


@pytest.mark.parametrize("stype", ["DWM", "PWM"])
def test_reference_pickle_loads_here(ref, tmp_path, stype):
    theirs = ref.atacorrect_functions.AtacBias(7, stype)
    _fill(theirs, stype, np.random.default_rng(1))
    path = str(tmp_path / "ref.pickle")
    theirs.to_pickle(path)
    mine = AtacBias().from_pickle(path)
    assert type(mine) is AtacBias and type(mine.bias["forward"]).__module__ == "tobias_b200.sequences"
    _same(mine, theirs, stype)


@pytest.mark.parametrize("stype", ["DWM", "PWM"])
def test_pickle_written_here_loads_in_the_reference(ref, tmp_path, stype):
    mine = AtacBias(7, stype)
    _fill(mine, stype, np.random.default_rng(2))
    path = str(tmp_path / "mine.pickle")
    mine.to_pickle(path)
    raw = open(path, "rb").read()
    assert b"tobias.tools.atacorrect_functions" in raw and b"tobias_b200" not in raw
    theirs = ref.atacorrect_functions.AtacBias().from_pickle(path)           # plain pickle.load inside the reference
    assert type(theirs).__module__ == "tobias.tools.atacorrect_functions"
    assert type(theirs.bias["reverse"]).__module__ == "tobias.utils.sequences"
    _same(mine, theirs, stype)
    # the reference can score with it right away
    seq = np.random.default_rng(3).integers(0, 4, size=60).astype(np.int64)
    sc = theirs.bias["forward"].score_sequence(seq)
    assert np.all(np.isfinite(sc[3:-3]))
    again = AtacBias().from_pickle(path)                                    # and this package reads its own files
    _same(again, mine, stype)
