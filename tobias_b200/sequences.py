"""
Host-side mirror of the reference's score-matrix classes (tobias/utils/sequences.pyx:23-342) for the hot path.

The per-read / per-base loops of the reference (`add_sequence`, `add_background`, `score_sequence`) run on the GPU
(K2 `tb_bias_counts`, K3a `tb_score_sequence`); what stays here is the bookkeeping the drivers do on the host: holding
the count tensors, summing them across chunks / ranks (`add_counts`, sequences.pyx:35-42) and turning counts into
log-odds tables (`prepare_mat`, :107-117 and :226-257 -- a few microseconds of numpy on 4 x 15 625 numbers, formulas
kept verbatim so that the tables are identical to the reference's).
"""
import sys

import numpy as np


class SequenceMatrix:
    @staticmethod
    def create(length, stype):
        if stype == "PWM":
            return NucleotideMatrix(length)
        if stype == "DWM":
            return DiNucleotideMatrix(length)
        sys.exit("Unkown score type {0}".format(stype))

    def add_counts(self, obj):
        self.counts += obj.counts
        self.bg_counts += obj.bg_counts
        self.neg_counts += obj.neg_counts
        self.no_bias += obj.no_bias
        self.no_bg += obj.no_bg

    # ---- array-level mirrors of the Cython methods (sequences.pyx:63-105, :178-223), computed through the C-ABI ----------------------
    def _add(self, sequence, amount, background, ctx):
        from . import signals
        ctx = ctx or signals._default_ctx()
        kmers = np.ascontiguousarray(sequence, dtype=np.int64).reshape(-1, self.length)
        amounts = np.broadcast_to(np.asarray(amount, dtype=np.float64), (kmers.shape[0],))
        target = self.bg_counts if background else self.counts
        target = np.ascontiguousarray(target, dtype=np.float64)
        neg = None if (background or self.stype == "DWM") else np.ascontiguousarray(self.neg_counts, dtype=np.float64)
        added = ctx.add_kmers(kmers, amounts, target, neg, is_dwm=self.stype == "DWM", is_background=background)
        if background:
            self.bg_counts = target
            self.no_bg += added
        else:
            self.counts = target
            if neg is not None:
                self.neg_counts = neg
            self.no_bias += added

    def add_sequence(self, sequence, amount=1.0, ctx=None):
        """One k-mer (int64[L] codes A0 T1 C2 G3 N4) `amount` times -- or a batch: [n, L] with `amount` a scalar or [n]."""
        self._add(sequence, amount, False, ctx)

    def add_background(self, sequence, amount=1.0, ctx=None):
        self._add(sequence, amount, True, ctx)

    # ---- device round trip -------------------------------------------------------------------------
    def upload(self, ctx, strand):
        """Hand the prepared tables of this strand (0 forward, 1 reverse) to the GPU context."""
        raise NotImplementedError

    def score_sequence(self, sequence, ctx=None):
        """Score a nucleotide-code array (A0 T1 C2 G3 N4) like the reference's score_sequence; runs on the GPU.
        `ctx` defaults to a scratch context on device 0."""
        from . import lib
        own = ctx is None
        if own:
            ctx = lib.Context(0)
        try:
            codes = np.asarray(sequence)
            ascii_arr = np.frombuffer(b"ATCGN", dtype=np.uint8)[np.minimum(codes, 4)]
            ctx.genome_load([ascii_arr])
            self.upload(ctx, 0)
            return ctx.score_sequence(0, 0, 0, codes.shape[0])
        finally:
            if own:
                ctx.close()


class NucleotideMatrix(SequenceMatrix):
    """PWM counts [5, L] (rows A,T,C,G,N) -- sequences.pyx:46-151."""
    stype = "PWM"

    def __init__(self, length):
        self.length = length
        self.counts = np.zeros((5, length))
        self.neg_counts = np.zeros((5, length))
        self.bg_counts = np.zeros((5, length))
        self.no_bias = 0
        self.no_bg = 0
        self.pssm = np.zeros((5, length))

    def prepare_mat(self):
        self.no_bias = np.mean(np.sum(self.counts[:4, :], axis=0))
        self.no_bg = np.mean(np.sum(self.bg_counts[:4, :], axis=0))
        self.bias_pwm = np.true_divide(self.counts + 1, self.no_bias + 4)
        self.bg_pwm = np.true_divide(self.bg_counts + 1, self.no_bg + 4)
        with np.errstate(divide="ignore", invalid="ignore"):
            self.pssm = np.log2(np.true_divide(self.bias_pwm, self.bg_pwm))

    def upload(self, ctx, strand):
        ctx.set_bias_model(strand, False, self.length, self.pssm)


class DiNucleotideMatrix(SequenceMatrix):
    """DWM counts [5, 5, L, L] (Sm, Sn, m, n) -- sequences.pyx:155-342."""
    stype = "DWM"

    def __init__(self, length):
        self.length = length
        self.counts = np.zeros((5, 5, length, length))
        self.neg_counts = np.zeros((5, length))
        self.bg_counts = np.zeros((5, 5, length, length))
        self.no_bias = 0
        self.no_bg = 0
        self.pssm = np.zeros((5, length))
        self.bias_pwm = np.zeros((5, 5, length, length))
        self.bg_pwm = np.zeros((5, 5, length, length))
        self.bias_dwm = np.zeros((5, 5, length, length))
        self.bg_dwm = np.zeros((5, 5, length, length))

    def prepare_mat(self):
        L = self.length
        bias_cm = np.sum(self.counts, axis=(0, 2)) / float(L)
        self.no_bias = np.mean(np.sum(bias_cm[:4, :], axis=0))
        self.bias_pwm = np.true_divide(bias_cm + 1, self.no_bias + 4)
        bg_cm = np.sum(self.bg_counts, axis=(0, 2)) / float(L)
        self.no_bg = np.mean(np.sum(bg_cm[:4, :], axis=0))
        self.bg_pwm = np.true_divide(bg_cm + 1, self.no_bg + 4)
        self.pssm = np.log2(np.true_divide(self.bias_pwm, self.bg_pwm))
        pseudo_bias = max([16, self.no_bias])
        pseudo_bg = max([16, self.no_bg])
        # dwm[Sm,Sn,m,n] = (counts[Sm,Sn,m,n] + pseudo * pwm[Sn,n] * pwm[Sm,m]) / (no + pseudo)   (sequences.pyx:246-251)
        self.bias_dwm = (self.counts + pseudo_bias * self.bias_pwm[None, :, None, :] * self.bias_pwm[:, None, :, None]) \
            / float(self.no_bias + pseudo_bias)
        self.bg_dwm = (self.bg_counts + pseudo_bg * self.bg_pwm[None, :, None, :] * self.bg_pwm[:, None, :, None]) \
            / float(self.no_bg + pseudo_bg)
        self.bias_pwm_log = np.log2(self.bias_pwm)
        self.bg_pwm_log = np.log2(self.bg_pwm)
        self.bias_dwm_log = np.log2(self.bias_dwm)
        self.bg_dwm_log = np.log2(self.bg_dwm)

    def upload(self, ctx, strand):
        ctx.set_bias_model(strand, True, self.length, self.bias_pwm_log, self.bg_pwm_log, self.bias_dwm_log, self.bg_dwm_log)
