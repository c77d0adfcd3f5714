"""
TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy + the C helpers in restate_c.c + scipy) of the TOBIAS
per-base hot path, working on the repo's columnar inputs.  It is the parity checker for the CUDA path;
the product (tobias_b200/) never imports it.  PINNED: tests/test_oracle_pinned.py checks every function
below against the compiled reference (oracle/_ref) and the fixtures in tests/golden/ were generated
by the reference itself (tests/golden/make_golden.py).

The one piece of third-party arithmetic on the path, scipy.optimize.curve_fit (MINPACK lmdif, called at
tobias/tools/atacorrect_functions.py:292), is called here exactly as the reference calls it
(scipy 1.18.1 in this image); upstream has no tests that pin results at that boundary.

All paths cited are relative to /root/reference.
"""
import ctypes
import os
import subprocess
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c(force=False):
    so = os.path.join(HERE, "librestate.so")
    src = os.path.join(HERE, "restate_c.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, src, "-lm"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_c())
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


_D, _U8 = ctypes.c_double, ctypes.c_uint8

# ------------------------------------------------------------------------------------------------ signals.pyx


def fast_rolling_math(arr, w, operation):
    """signals.pyx:102-177 ("sum" / "mean" only -- the operations the hot path uses)."""
    arr = np.ascontiguousarray(arr, dtype=np.float64)
    out = np.empty_like(arr)
    _lib().o_rolling(_p(arr, _D), arr.shape[0], int(w), {"sum": 0, "mean": 1}[operation], _p(out, _D))
    return out


def tobias_footprint_array(arr, flank_min, flank_max, fp_min, fp_max):
    """signals.pyx:184-230."""
    arr = np.ascontiguousarray(arr, dtype=np.float64)
    out = np.empty_like(arr)
    _lib().o_footprint(_p(arr, _D), arr.shape[0], flank_min, flank_max, fp_min, fp_max, _p(out, _D))
    return out


def FOS_score(arr, flank_min, flank_max, fp_min, fp_max):
    """signals.pyx:237-282."""
    arr = np.ascontiguousarray(arr, dtype=np.float64)
    out = np.empty_like(arr)
    _lib().o_fos(_p(arr, _D), arr.shape[0], flank_min, flank_max, fp_min, fp_max, _p(out, _D))
    return out


def calculate_scores_region(signal_f32, flank, score="footprint", absolute=False, min_limit=None, max_limit=None,
                            window=100, smooth=1, flank_min=10, flank_max=30, fp_min=20, fp_max=50):
    """tools/score_bigwig.py:53-94 for one region; `signal_f32` = bigwig values of the region extended by `flank`."""
    signal = np.nan_to_num(np.asarray(signal_f32)).astype("float64")
    if absolute:
        signal = np.abs(signal)
    if min_limit is not None:
        signal[signal < min_limit] = min_limit
    if max_limit is not None:
        signal[signal > max_limit] = max_limit
    if score == "sum":
        scores = fast_rolling_math(signal, window, "sum")
    elif score == "mean":
        scores = fast_rolling_math(signal, window, "mean")
    elif score == "footprint":
        scores = tobias_footprint_array(signal, flank_min, flank_max, fp_min, fp_max)
    elif score == "FOS":
        scores = FOS_score(signal, flank_min, flank_max, fp_min, fp_max)
    elif score == "none":
        scores = signal
    else:
        raise ValueError(score)
    if smooth > 1:
        scores = fast_rolling_math(scores, smooth, "mean")
    if flank > 0:
        scores = scores[flank:-flank]
    return scores


# ------------------------------------------------------------------------------------------------ sequences.pyx
_CODE = np.full(256, 4, dtype=np.uint8)
for _i, _c in enumerate("ATCG"):          # sequences.pyx:398-408
    _CODE[ord(_c)] = _i
    _CODE[ord(_c.lower())] = _i
_COMP = np.array([1, 0, 3, 2, 4], dtype=np.uint8)     # sequences.pyx:411-420


def encode(ascii_arr):
    return _CODE[np.asarray(ascii_arr, dtype=np.uint8)]


def revcomp(codes):
    return _COMP[codes][::-1]


class Matrix:
    """Count + score tables of one strand: NucleotideMatrix (sequences.pyx:46-151) or DiNucleotideMatrix (:155-342)."""

    def __init__(self, length, stype):
        self.length, self.stype = length, stype
        shape = (5, 5, length, length) if stype == "DWM" else (5, length)
        self.counts = np.zeros(shape)
        self.bg_counts = np.zeros(shape)
        self.neg_counts = np.zeros((5, length))
        self.no_bias = 0
        self.no_bg = 0

    def add_sequence(self, kmer, amount=1.0):
        L = self.length
        if self.stype == "DWM":                                   # :178-198
            k = np.ascontiguousarray(kmer, dtype=np.uint8)
            if _lib().o_dwm_add(_p(self.counts, _D), _p(k, _U8), L, ctypes.c_double(amount)):
                self.no_bias += amount
        else:                                                     # :63-88
            if np.any(np.asarray(kmer) == 4):
                return
            tgt = self.counts if amount > 0 else self.neg_counts
            tgt[np.asarray(kmer), np.arange(L)] += amount
            self.no_bias += amount

    def add_background(self, kmer, amount=1.0):
        L = self.length
        if self.stype == "DWM":                                   # :203-223
            k = np.ascontiguousarray(kmer, dtype=np.uint8)
            if _lib().o_dwm_add(_p(self.bg_counts, _D), _p(k, _U8), L, ctypes.c_double(amount)):
                self.no_bg += amount
        else:                                                     # :93-105 -- no N check in the PWM variant
            self.bg_counts[np.asarray(kmer), np.arange(L)] += amount
            self.no_bg += amount

    def add_counts(self, other):                                  # :35-42
        self.counts += other.counts
        self.bg_counts += other.bg_counts
        self.neg_counts += other.neg_counts
        self.no_bias += other.no_bias
        self.no_bg += other.no_bg

    def prepare_mat(self):
        L = self.length
        if self.stype == "PWM":                                   # :107-117
            self.no_bias = np.mean(np.sum(self.counts[:4, :], axis=0))
            self.no_bg = np.mean(np.sum(self.bg_counts[:4, :], axis=0))
            self.bias_pwm = np.true_divide(self.counts + 1, self.no_bias + 4)
            self.bg_pwm = np.true_divide(self.bg_counts + 1, self.no_bg + 4)
            with np.errstate(all="ignore"):
                self.pssm = np.log2(np.true_divide(self.bias_pwm, self.bg_pwm))
            return
        bias_cm = np.sum(self.counts, axis=(0, 2)) / float(L)    # :231-241
        self.no_bias = np.mean(np.sum(bias_cm[:4, :], axis=0))
        self.bias_pwm = np.true_divide(bias_cm + 1, self.no_bias + 4)
        bg_cm = np.sum(self.bg_counts, axis=(0, 2)) / float(L)
        self.no_bg = np.mean(np.sum(bg_cm[:4, :], axis=0))
        self.bg_pwm = np.true_divide(bg_cm + 1, self.no_bg + 4)
        self.pssm = np.log2(np.true_divide(self.bias_pwm, self.bg_pwm))
        pseudo_bias = max([16, self.no_bias])                     # :244-251
        pseudo_bg = max([16, self.no_bg])
        self.bias_dwm = (self.counts + pseudo_bias * self.bias_pwm[None, :, None, :] * self.bias_pwm[:, None, :, None]) \
            / float(self.no_bias + pseudo_bias)
        self.bg_dwm = (self.bg_counts + pseudo_bg * self.bg_pwm[None, :, None, :] * self.bg_pwm[:, None, :, None]) \
            / float(self.no_bg + pseudo_bg)
        self.bias_pwm_log = np.log2(self.bias_pwm)                # :254-257
        self.bg_pwm_log = np.log2(self.bg_pwm)
        self.bias_dwm_log = np.log2(self.bias_dwm)
        self.bg_dwm_log = np.log2(self.bg_dwm)

    def score_sequence(self, codes):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        out = np.empty(codes.shape[0])
        L = self.length
        if self.stype == "PWM":
            pssm = np.ascontiguousarray(self.pssm)
            _lib().o_score_pwm(_p(codes, _U8), codes.shape[0], _p(pssm, _D), L, _p(out, _D))
        else:
            t = [np.ascontiguousarray(x) for x in (self.bias_pwm_log, self.bg_pwm_log, self.bias_dwm_log, self.bg_dwm_log)]
            _lib().o_score_dwm(_p(codes, _U8), codes.shape[0], _p(t[0], _D), _p(t[1], _D), _p(t[2], _D), _p(t[3], _D),
                               L, _p(out, _D))
        return out


class AtacBias:
    """tools/atacorrect_functions.py:43-60 (forward / reverse matrices + no_reads)."""

    def __init__(self, L=10, stype="DWM"):
        self.stype = stype
        self.bias = {"forward": Matrix(L, stype), "reverse": Matrix(L, stype)}
        self.no_reads = 0
        self.correction_factor = 1.0

    def join(self, other):
        for s in ("forward", "reverse"):
            self.bias[s].add_counts(other.bias[s])
        self.no_reads += other.no_reads


# ------------------------------------------------------------------------------------------------ ngs.pyx


def cutsites(reads, read_shift=(4, -5)):
    """1-based cut site of every record, OneRead.get_cutsite (ngs.pyx:91-106)."""
    rev = reads.is_reverse
    fwd_pos = reads.ref_start.astype(np.int64) - reads.left_soft                       # reference_start - query_alignment_start
    rev_pos = reads.ref_end.astype(np.int64) + reads.right_soft                       # reference_end + query_length - query_alignment_end
    return np.where(rev, rev_pos + 1 + read_shift[1], fwd_pos + 1 + read_shift[0])


def fetch(reads, chrom_id, start, end):
    """ReadList.from_bam (ngs.pyx:150-160): overlapping records minus unmapped / duplicate."""
    idx = reads.fetch_index(chrom_id, start, end)
    if idx.size:
        idx = idx[reads.usable()[idx]]
    return idx


def count_reads(reads, regions, read_shift=(4, -5)):
    """tools/atacorrect_functions.py:80-105; regions = iterable of (chrom_id, start, end)."""
    cuts = cutsites(reads, read_shift)
    total = 0
    for cid, start, end in regions:
        idx = fetch(reads, cid, start, end)
        c = cuts[idx]
        total += int(np.count_nonzero((c > start) & (c <= end)))                      # :99
    return total


def pileup(reads, chrom_id, start, end, read_shift=(4, -5)):
    """Per-strand cut-site counts of region [start, end): atacorrect_functions.py:231-243 + ReadList.signal
    (ngs.pyx:186-199).  Returns (forward, reverse) float64 arrays."""
    idx = fetch(reads, chrom_id, start, end)
    c = cutsites(reads, read_shift)[idx]
    keep = (c > start) & (c < end)                                                    # :237 (strict on both sides)
    idx, c = idx[keep], c[keep]
    rev = reads.is_reverse[idx]
    out = []
    for strand_rev in (False, True):
        v = np.zeros(end - start)
        np.add.at(v, (c[rev == strand_rev] - start - 1), 1.0)
        out.append(np.round(v, 5))                                                    # :243
    return out[0], out[1]


def _kmer(codes_ext, rc_ext, seq_start, seq_end, cutsite, is_reverse, k_flank):
    """OneRead.get_kmer (ngs.pyx:109-128) on a GenomicSequence of [seq_start, seq_end)."""
    if cutsite > seq_start + k_flank + 1 and cutsite <= seq_end - k_flank:
        if is_reverse:
            i = seq_end - cutsite
            return rc_ext[i - k_flank:i + k_flank + 1]
        i = cutsite - 1 - seq_start
        return codes_ext[i - k_flank:i + k_flank + 1]
    return np.full(2 * k_flank + 1, 4, dtype=np.uint8)


def bias_estimation(reads, genome, regions, chrom_lengths, k_flank=12, bg_shift=100, read_shift=(4, -5),
                    score_mat="DWM", cap=10):
    """tools/atacorrect_functions.py:109-181.  regions = iterable of (chrom_id, start, end); genome = list of ASCII
    arrays by chrom_id.  QUIRK kept: the 'within borders' test (:166) uses the region AFTER it was extended in
    place by k_flank+bg_shift and clipped (:142-143)."""
    L = 2 * k_flank + 1
    bias_obj = AtacBias(L, score_mat)
    cuts_all = cutsites(reads, read_shift)
    ok5 = reads.five_prime_is_match()
    for cid, start, end in regions:
        idx = fetch(reads, cid, start, end)                                           # :134 (original bounds)
        if idx.size == 0:                                                             # :139
            continue
        ext_start = max(0, start - (k_flank + bg_shift))                              # :142-143
        ext_end = min(int(chrom_lengths[cid]), end + (k_flank + bg_shift))
        codes = encode(genome[cid][ext_start:ext_end])                                # :144
        rc = revcomp(codes)
        rev = reads.is_reverse[idx]
        for strand, is_rev in (("forward", False), ("reverse", True)):
            sel = idx[(rev == is_rev) & ok5[idx]]                                     # :155-159
            if sel.size == 0:
                continue
            uniq, n = np.unique(cuts_all[sel], return_counts=True)
            mat = bias_obj.bias[strand]
            for c, cnt in zip(uniq.tolist(), n.tolist()):
                if c > ext_start and c < ext_end:                                     # :166 on the mutated region
                    w = min(cnt, cap)                                                 # :168
                    mat.add_sequence(_kmer(codes, rc, ext_start, ext_end, c, is_rev, k_flank), w)     # :170-172
                    cb = c + bg_shift if is_rev else c - bg_shift                     # :173 shift_cutsite(-bg_shift)
                    mat.add_background(_kmer(codes, rc, ext_start, ext_end, cb, is_rev, k_flank), w)  # :174-175
                    bias_obj.no_reads += w                                            # :177
    return bias_obj


# ------------------------------------------------------------------------------------------------ bias_correction


def relu(x, a, b):
    """tools/atacorrect_functions.py:185-188."""
    return np.maximum(0.0, a * x + b)


def fit_window(signal_w, bias_w, trace=None):
    """One 100-bp window of atacorrect_functions.py:285-304: returns bias_predict (normalised)."""
    from scipy.optimize import curve_fit, OptimizeWarning
    if np.max(signal_w) > 0:
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("error", OptimizeWarning)                       # :37-40
                popt, pcov, info, msg, ier = curve_fit(relu, bias_w, signal_w, full_output=True)
            pred = relu(bias_w, *popt)
            if trace is not None:
                trace.append((popt[0], popt[1], ier, info["nfev"], 0))
        except (OptimizeWarning, RuntimeError):                                       # :295-299
            cut_positions = np.logical_not(np.isclose(signal_w, 0))
            bias_min = np.min(bias_w[cut_positions])
            pred = bias_w - bias_min
            pred[pred < 0] = 0
            if trace is not None:
                trace.append((np.nan, bias_min, -1, 0, 1))
        if np.max(pred) > 0:                                                          # :301-302
            pred = pred / np.max(pred)
    else:
        pred = np.zeros(signal_w.shape[0])                                            # :304
        if trace is not None:
            trace.append((np.nan, np.nan, 0, 0, 2))
    return pred


def correct_strand(signal, bias, k_flank, window, correction_factor, trace=None):
    """Per-strand body of atacorrect_functions.py:269-350.  signal = rounded pileup, bias = nan_to_num(score).
    Returns (uncorrected*cf, bias_prediction, expected*cf, corrected)."""
    reg_len = signal.shape[0]
    w = window
    reg_end = reg_len - k_flank
    step = 10
    overlaps = int(window / step)
    window_starts = list(range(k_flank, reg_end - window, step))                      # :272-274
    window_ends = list(range(k_flank + window, reg_end, step))
    windows = list(zip(window_starts, window_ends))
    bias_predictions = np.zeros((overlaps, reg_len))
    bias_predictions[k_flank:reg_end] = np.nan                                       # :280 (indexes ROWS -- QUIRK)
    row = 0
    bias_prediction = None
    for ws, we in windows:
        bias_predictions[row, ws:we] = fit_window(signal[ws:we], bias[ws:we], trace)
        row = row + 1 if row < overlaps - 1 else 0
    if windows:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            bias_prediction = np.nanmean(bias_predictions, axis=0)                    # :309
    else:
        raise ValueError("region too short for a single window (reference behaviour undefined, :311)")
    b = bias_prediction
    signal_sum = fast_rolling_math(signal, w, "sum")                                  # :314-321
    signal_sum[np.isnan(signal_sum)] = 0
    bias_sum = fast_rolling_math(b, w, "sum")
    nulls = np.logical_or(np.isclose(bias_sum, 0), np.isnan(bias_sum))
    bias_sum[nulls] = 1
    bias_probas = b / bias_sum
    bias_probas[nulls] = 0
    expected = signal_sum * bias_probas                                               # :323
    unc = signal * correction_factor                                                  # :326-328
    expected = expected * correction_factor
    corrected = unc - expected
    uncorrected_sum = fast_rolling_math(unc, w, "sum")                                # :331-350
    uncorrected_sum[np.isnan(uncorrected_sum)] = 0
    corrected_sum = fast_rolling_math(np.abs(corrected), w, "sum")
    corrected_sum[np.isnan(corrected_sum)] = 0
    corrected_pos = np.copy(corrected)
    corrected_pos[corrected_pos < 0] = 0
    corrected_pos_sum = fast_rolling_math(corrected_pos, w, "sum")
    corrected_pos_sum[np.isnan(corrected_pos_sum)] = 0
    corrected_neg_sum = corrected_sum - corrected_pos_sum
    zero_sum = corrected_pos_sum == 0
    corrected_pos_sum[zero_sum] = np.nan
    with np.errstate(all="ignore"):
        scale_factor = (uncorrected_sum - corrected_neg_sum) / corrected_pos_sum
    scale_factor[zero_sum] = 1
    scale_factor[scale_factor < 1] = 1
    pos_bool = corrected > 0
    corrected[pos_bool] *= scale_factor[pos_bool]
    return unc, b, expected, corrected


def bias_correction_region(reads, genome, chrom_id, start, end, bias_obj, k_flank=12, window=100,
                           read_shift=(4, -5), trace=None):
    """One output region [start, end) of tools/atacorrect_functions.py:218-400 (before the writer queue).
    Returns dict track -> {forward, reverse} trimmed float64 arrays, and the pre/post verification PWM counts
    as (counts[5,L], neg_counts[5,L]) per strand."""
    L = 2 * k_flank + 1
    f = int(window / 2.0)
    f_extend = k_flank + f
    s, e = start - f_extend, end + f_extend                                           # :220
    reg_len = e - s
    unc = dict(zip(("forward", "reverse"), pileup(reads, chrom_id, s, e, read_shift)))
    codes = encode(genome[chrom_id][s:e])                                             # :251
    rc = revcomp(codes)
    out = {t: {} for t in ("uncorrected", "bias", "expected", "corrected")}
    bias = {}
    bias["forward"] = np.nan_to_num(bias_obj.bias["forward"].score_sequence(codes))   # :254-262
    bias["reverse"] = np.nan_to_num(bias_obj.bias["reverse"].score_sequence(rc)[::-1])
    pre = {st: Matrix(L, "PWM") for st in ("forward", "reverse")}
    post = {st: Matrix(L, "PWM") for st in ("forward", "reverse")}
    for strand in ("forward", "reverse"):
        u, b, ex, co = correct_strand(unc[strand], bias[strand], k_flank, window, bias_obj.correction_factor,
                                      None if trace is None else trace.setdefault(strand, []))
        out["uncorrected"][strand], out["expected"][strand], out["corrected"][strand] = u, ex, co
        out["bias"][strand] = bias[strand]
        for idx in range(k_flank + 1, reg_len - k_flank - 1):                         # :359-360
            orig, corr = u[idx], co[idx]
            if orig != 0 or corr != 0:
                if strand == "forward":
                    kmer = codes[idx - k_flank:idx + k_flank + 1]
                else:
                    kmer = rc[reg_len - idx - k_flank - 1:reg_len - idx + k_flank]
                pre[strand].add_sequence(kmer, orig)
                post[strand].add_sequence(kmer, corr)
    for t in out:                                                                     # :381-383
        for st in out[t]:
            out[t][st] = out[t][st][f_extend:-f_extend]
    return out, pre, post
