/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C) of the native kernels of the TOBIAS per-base
 * hot path.  Used as the parity checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg; never linked into or called by the product (tobias_b200/).
 *
 * Each function cites the reference lines it restates (paths relative to /root/reference).  The
 * restatement is pinned against the compiled reference itself (oracle/_ref) by
 * tests/test_oracle_pinned.py and against the committed fixtures under tests/golden/.
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -o oracle/librestate.so oracle/restate_c.c -lm
 * (-ffp-contract=off: the reference is compiled without FMA contraction on x86-64.)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* ---- fast_rolling_math, tobias/utils/signals.pyx:102-177 ----------------------------------------
 * op 0 = sum (:149-155), 1 = mean (:158-164).  The accumulator is a C float (`cdef float valsum`,
 * :116): every `valsum += arr[i+j]` is a double addition rounded to float.  Flanks are NaN (:112). */
void o_rolling(const double *arr, int L, int w, int op, double *out) {
    int lf = (int)floor(w / 2.0);
    for (int i = 0; i < L; i++) out[i] = NAN;
    for (int i = 0; i < L - w + 1; i++) {
        float valsum = 0;
        for (int j = 0; j < w; j++) valsum = (float)((double)valsum + arr[i + j]);
        out[i + lf] = (op == 0) ? (double)valsum : (double)valsum / (w * 1.0);
    }
}

/* ---- tobias_footprint_array, tobias/utils/signals.pyx:184-230 -----------------------------------
 * all intermediates are C floats (:190-191); both width ranges are inclusive (:195-196). */
void o_footprint(const double *arr, int L, int flank_min, int flank_max, int fp_min, int fp_max, double *out) {
    for (int i = 0; i < L; i++) out[i] = 0.0;
    for (int i = 0; i < L - 2 * flank_max - fp_max; i++) {
        for (int fw = flank_min; fw <= flank_max; fw++) {
            for (int pw = fp_min; pw <= fp_max; pw++) {
                float left = 0.0f, mid = 0.0f, right = 0.0f, val;
                for (int j = 0; j < fw; j++) { val = (float)arr[i + j]; if (val > 0.0) left += val; }
                for (int j = 0; j < pw; j++) { val = (float)arr[i + fw + j]; if (val < 0.0) mid += val; }
                for (int j = 0; j < fw; j++) { val = (float)arr[i + fw + pw + j]; if (val > 0.0) right += val; }
                float fp_mean = (float)(mid / (1.0 * pw));
                float flank_mean = (float)((right + left) / (2.0 * fw));
                float score = flank_mean - fp_mean;
                for (int pos = i + fw; pos < i + fw + pw; pos++)
                    if (score > out[pos]) out[pos] = score;
            }
        }
    }
}

/* ---- FOS_score, tobias/utils/signals.pyx:237-282 ------------------------------------------------
 * width ranges are EXCLUSIVE here (:249-250); running minimum, initial value 10000 (:241). */
void o_fos(const double *arr, int L, int flank_min, int flank_max, int fp_min, int fp_max, double *out) {
    for (int i = 0; i < L; i++) out[i] = 10000.0;
    for (int i = 0; i < L - 2 * flank_max - fp_max; i++) {
        for (int fw = flank_min; fw < flank_max; fw++) {
            for (int pw = fp_min; pw < fp_max; pw++) {
                float left = 0.0f, mid = 0.0f, right = 0.0f;
                for (int j = 0; j < fw; j++) left = (float)((double)left + arr[i + j]);
                for (int j = 0; j < pw; j++) mid = (float)((double)mid + arr[i + fw + j]);
                for (int j = 0; j < fw; j++) right = (float)((double)right + arr[i + fw + pw + j]);
                float Lm = (float)(left / (fw * 1.0));
                float Cm = (float)(mid / (pw * 1.0));
                float Rm = (float)(right / (fw * 1.0));
                float score;
                if (Cm < Rm && Cm < Lm && Lm > 0 && Rm > 0) score = (float)((Cm + 1.0) / Lm + (Cm + 1.0) / Rm);
                else score = 10000;
                for (int j = 0; j < pw; j++)
                    if (score < out[i + fw + j]) out[i + fw + j] = score;
            }
        }
    }
}

/* ---- NucleotideMatrix.score_sequence, tobias/utils/sequences.pyx:123-151 ------------------------
 * pssm is [5][L] row-major (rows A,T,C,G,N).  Windows containing N stay NaN. */
void o_score_pwm(const uint8_t *seq, int n, const double *pssm, int L, double *out) {
    int F = (int)(L / 2.0);
    for (int i = 0; i < n; i++) out[i] = NAN;
    for (int i = 0; i < n - L + 1; i++) {
        int has_n = 0;
        double score = 0.0;
        for (int m = 0; m < L; m++) {
            int s = seq[i + m];
            if (s > 3) has_n = 1;
            score += pssm[s * L + m];
        }
        if (!has_n) out[i + F] = score;
    }
}

/* ---- DiNucleotideMatrix.score_sequence, tobias/utils/sequences.pyx:263-342 ----------------------
 * pwm tables are [5][L]; dwm tables are [5][5][L][L] (Sm, Sn, m, n), all log2.  Same loop order and
 * the same floating-point operation order as the reference (so results agree to the last bit with
 * a build that does not contract multiply-adds). */
#define DW(t, sm, sn, m, n) ((t)[(((size_t)(sm) * 5 + (sn)) * L + (m)) * L + (n)])
static double dwm_half(const uint8_t *k, const double *pwm, const double *dwm, int L) {
    double total = 0;
    for (int n = 0; n < L; n++) {
        int sn = k[n];
        double cn = 0.0;
        for (int a = 0; a < 4; a++) {
            double prod = 0.0;
            for (int m = 0; m < L; m++)
                if (m != n) prod += DW(dwm, k[m], a, m, n) - pwm[a * L + n];
            cn += pow(2.0, prod + pwm[a * L + n]);
        }
        double cond = 0.0;
        for (int m = 0; m < L; m++)
            if (m != n) cond += DW(dwm, k[m], sn, m, n) - pwm[sn * L + n];
        total += cond + pwm[sn * L + n] - log2(cn);
    }
    return total;
}
void o_score_dwm(const uint8_t *seq, int n, const double *bias_pwm, const double *bg_pwm,
                 const double *bias_dwm, const double *bg_dwm, int L, double *out) {
    int F = (int)(L / 2.0);
    for (int i = 0; i < n; i++) out[i] = NAN;
    for (int i = 0; i < n - L + 1; i++) {
        int has_n = 0;
        for (int m = 0; m < L; m++) if (seq[i + m] > 3) has_n = 1;
        if (has_n) continue;
        /* the reference interleaves bg and bias per position n; the two running sums are independent */
        double bg = dwm_half(seq + i, bg_pwm, bg_dwm, L);
        double bias = dwm_half(seq + i, bias_pwm, bias_dwm, L);
        out[i + F] = bias - bg;
    }
}

/* ---- DiNucleotideMatrix.add_sequence / add_background, tobias/utils/sequences.pyx:178-223 -------
 * counts is [5][5][L][L]; returns 0 when the k-mer holds an N (nothing added). */
int o_dwm_add(double *counts, const uint8_t *kmer, int L, double amount) {
    for (int m = 0; m < L; m++) if (kmer[m] == 4) return 0;
    for (int m = 0; m < L; m++)
        for (int n = 0; n < L; n++)
            DW(counts, kmer[m], kmer[n], m, n) += amount;
    return 1;
}
