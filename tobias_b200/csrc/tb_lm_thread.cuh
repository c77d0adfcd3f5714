// tb_lm_thread.cuh -- K3b core, thread-per-window form: ONE THREAD fits  signal ~ max(0, a*bias + b)  on one window.
//
// Same algorithm and same decisions as tb_lm.cuh (scipy.optimize.curve_fit -> leastsq -> MINPACK lmdif with scipy's
// settings, called by the reference at tobias/tools/atacorrect_functions.py:292), restated so that a thread needs NO
// m-vector storage: MINPACK keeps fvec, wa4 and the m x 2 Jacobian / Householder vectors in memory; here every
// element of them is recomputed from the window's (bias, signal) pair whenever a pass over the window needs it, with
// exactly the operations MINPACK applied to it, so each element has the same bits every time it is formed.  The
// m-term sums run in index order in one thread -- MINPACK's own order -- so the fit reproduces scipy bit for bit
// (tests/test_lm_exact.py) while the 32 lanes of a warp work on 32 different windows (no shuffles, no redundant
// trust-region algebra).
//
// Passes over the window per outer iteration: A column norms of the forward-difference Jacobian (fdjac2 + qrfac's
// enorm), B Householder column 0 applied to column 1 and to fvec, C norm of the remaining column, D second component
// of Q^T fvec; per inner iteration one pass for the trial residual norm.
//
// Divisions by a pass-invariant divisor d (h of fdjac2, the Householder norms) use the precomputed IEEE reciprocal and
// one fused correction step (q = a*(1/d); r = fma(-q, d, a); q + r*(1/d)), which returns the correctly rounded
// quotient a/d (Markstein's theorem; checked against the divide instruction on 1.6e9 operand pairs).
#pragma once
#include "tb_common.cuh"
#include "tb_lm.cuh"

// Window point i = 10 u + v of this thread's window lives at x[v * stride + u] (bias, f64) / y[v * stride + u]
// (signal, u32 counts): consecutive windows start 10 positions apart, so consecutive windows read consecutive slots.
struct tb_win_ref {
    const double *x;
    const u32 *y;
    i64 stride;             // distance between the planes v = 0..9
};
// exact u32 -> double without a conversion instruction: (2^52 + k) - 2^52
__device__ __forceinline__ double tb_u32_to_double(u32 k) {
    return __hiloint2double(0x43300000, (int)k) - 4503599627370496.0;
}

// for (i = i0; i < m; i++) body(i, ex = bias[i], ey = signal[i] as double, eyi = signal[i] as u32), i0 < 10.  Kept compact
// on purpose (no unrolling): the kernel is a sequence of such passes and must stay resident in the instruction cache.
// Point i + 1 is loaded while point i is processed (hides the L1 latency); offsets are 32-bit element indices
// (the host checks 10 * stride < 2^32).
#ifndef TB_FIT_UNROLL
#define TB_FIT_UNROLL 1
#endif
#define TB_PRAGMA_(x) _Pragma(#x)
#define TB_PRAGMA(x) TB_PRAGMA_(x)
#define TB_FOR_POINTS_FROM(W, m, i0, i, ...)                                   \
    {                                                                          \
        const u32 _st = (u32)(W).stride, _wrap = 9u * _st - 1u;                \
        u32 _off = (u32)(i0) * _st;                                            \
        double _nx = (W).x[_off];                                              \
        u32 _ny = (W).y[_off];                                                 \
        int _v = (i0);                                                         \
        TB_PRAGMA(unroll TB_FIT_UNROLL)                                        \
        for (int i = (i0); i < (m); i++) {                                     \
            const double ex = _nx;                                             \
            const u32 eyi = _ny;                                               \
            const bool _w10 = ++_v == 10;                                      \
            _off = _w10 ? _off - _wrap : _off + _st;                           \
            _v = _w10 ? 0 : _v;                                                \
            if (i + 1 < (m)) {                                                 \
                _nx = (W).x[_off];                                             \
                _ny = (W).y[_off];                                             \
            }                                                                  \
            const double ey = tb_u32_to_double(eyi);                           \
            __VA_ARGS__                                                        \
        }                                                                      \
    }
#define TB_FOR_POINTS(W, m, i, ...) TB_FOR_POINTS_FROM(W, m, 0, i, __VA_ARGS__)
// all points in storage order (plane by plane: v = 0..9, then u): for sums whose order is free (the one-pass form).  No index
// arithmetic per point -- consecutive points of a plane are consecutive slots.
// One flat loop over the m points (plane by plane) with the loads running TB_FIT_AHEAD points ahead of the arithmetic, across plane
// boundaries.  Measured on hg38 (K3b): 2 ahead 59.8 ms, 3: 63.9, 4: 66.9, 6: 67.4 (register pressure: the round kernels sit at the
// 128-register limit of four blocks per SM); the per-plane loop with one point ahead it replaces: 61.0.
#ifndef TB_FIT_AHEAD
#define TB_FIT_AHEAD 2
#endif
#define TB_FOR_POINTS_ANY_ORDER(W, m, ...)                                                                  \
    {                                                                                                       \
        const double *_px = (W).x;                                                                          \
        const u32 *_py = (W).y;                                                                             \
        int _v = 0, _left = ((m) + 9) / 10;            /* points of plane 0 still to load */                \
        double _qx[TB_FIT_AHEAD];                                                                           \
        u32 _qy[TB_FIT_AHEAD];                                                                              \
        _Pragma("unroll") for (int _k = 0; _k < TB_FIT_AHEAD; _k++) {                                       \
            _qx[_k] = 0.0;                                                                                  \
            _qy[_k] = 0u;                                                                                   \
            if (_k < (m)) {                                                                                 \
                _qx[_k] = *_px;                                                                             \
                _qy[_k] = *_py;                                                                             \
                if (--_left == 0) {                                                                         \
                    _v++;                                                                                   \
                    _left = ((m) - _v + 9) / 10;                                                            \
                    _px = (W).x + (i64)_v * (W).stride;                                                     \
                    _py = (W).y + (i64)_v * (W).stride;                                                     \
                } else {                                                                                    \
                    _px++;                                                                                  \
                    _py++;                                                                                  \
                }                                                                                           \
            }                                                                                               \
        }                                                                                                   \
        for (int _i0 = 0; _i0 < (m); _i0 += TB_FIT_AHEAD) {                                                 \
            _Pragma("unroll") for (int _k = 0; _k < TB_FIT_AHEAD; _k++) {                                   \
                if (_i0 + _k < (m)) {                                                                       \
                    const double ex = _qx[_k];                                                              \
                    const u32 eyi = _qy[_k];                                                                \
                    if (_i0 + _k + TB_FIT_AHEAD < (m)) {                                                    \
                        _qx[_k] = *_px;                                                                     \
                        _qy[_k] = *_py;                                                                     \
                        if (--_left == 0) {                                                                 \
                            _v++;                                                                           \
                            _left = ((m) - _v + 9) / 10;                                                    \
                            _px = (W).x + (i64)_v * (W).stride;                                             \
                            _py = (W).y + (i64)_v * (W).stride;                                             \
                        } else {                                                                            \
                            _px++;                                                                          \
                            _py++;                                                                          \
                        }                                                                                   \
                    }                                                                                       \
                    const double ey = tb_u32_to_double(eyi);                                                \
                    __VA_ARGS__                                                                             \
                }                                                                                           \
            }                                                                                               \
        }                                                                                                   \
    }

// a / d with dinv = 1.0 / d (correctly rounded, see header)
__device__ __forceinline__ double tb_div_inv(double a, double d, double dinv) {
    double q = a * dinv;
    double r = fma(-q, d, a);
    return fma(r, dinv, q);
}

// MINPACK enorm as a streaming accumulator (three sums: large / intermediate / small components).  A component whose
// exponent is surely inside (3.834e-20, 1.304e19 / n) is squared and added in line (integer test of the high word);
// exact zeros are skipped (they change nothing in MINPACK's small branch); anything else takes the out-of-line path
// with MINPACK's full classification, which practically never runs.
struct tb_enorm_acc {
    double s1, s2, s3, x1max, x3max;
};
__device__ __noinline__ tb_enorm_acc tb_enorm_rare(tb_enorm_acc t, double xabs, double agiant) {
    if (xabs > 3.834e-20 && xabs < agiant) {
        t.s2 += xabs * xabs;
    } else if (xabs <= 3.834e-20) {
        if (xabs > t.x3max) {
            double q = t.x3max / xabs;
            t.s3 = 1.0 + t.s3 * (q * q);
            t.x3max = xabs;
        } else if (xabs != 0.0) {
            double q = xabs / t.x3max;
            t.s3 += q * q;
        }
    } else {
        if (xabs > t.x1max) {
            double q = t.x1max / xabs;
            t.s1 = 1.0 + t.s1 * (q * q);
            t.x1max = xabs;
        } else {
            double q = xabs / t.x1max;
            t.s1 += q * q;
        }
    }
    return t;
}
__device__ __noinline__ double tb_enorm_result(tb_enorm_acc t) {
    if (t.s1 != 0.0) return t.x1max * sqrt(t.s1 + (t.s2 / t.x1max) / t.x1max);
    if (t.s2 != 0.0) {
        if (t.s2 >= t.x3max) return sqrt(t.s2 * (1.0 + (t.x3max / t.s2) * (t.x3max * t.s3)));
        return sqrt(t.x3max * ((t.s2 / t.x3max) + (t.x3max * t.s3)));
    }
    return t.x3max * sqrt(t.s3);
}
__device__ __forceinline__ void tb_enorm_init(tb_enorm_acc &t) { t.s1 = t.s2 = t.s3 = t.x1max = t.x3max = 0.0; }
// n = number of components of the whole vector (sets MINPACK's agiant); n <= 512 keeps the in-line range test valid
__device__ __forceinline__ void tb_enorm_add(tb_enorm_acc &t, double v, int n) {
    const u32 hi = (u32)__double2hiint(v) & 0x7fffffffu;
    if (hi - 0x3BF00000u < 0x43500000u - 0x3BF00000u) t.s2 += v * v;             // 2^-64 <= |v| < 2^54
    else if ((hi | (u32)__double2loint(v)) != 0u) t = tb_enorm_rare(t, fabs(v), 1.304e19 / (double)n);
}

// state of one fit between passes
struct tb_lmt {
    double par0, par1, fnorm, diag0, diag1, delta, lmp, xnorm, gnorm;
    double r00, r01, r11, qtf0, qtf1, acn0, acn1;
    double tyy;                                         // moment form: sum of signal^2 over the window (a constant of the fit)
    int ipvt0, iter, nfev, info;
};

// max(0, t) as the reference's np.maximum(0.0, t) for every finite t (and -0.0 -> 0.0): clears negative values by their
// sign bit on the integer pipe instead of a double-precision compare + select
__device__ __forceinline__ double tb_relu(double t) {
    const int hi = __double2hiint(t);
    const int keep = ~(hi >> 31);                                // all ones for a clear sign bit
    return __hiloint2double(hi & keep, __double2loint(t) & keep);
}
__device__ __forceinline__ double tb_relu_res(double a, double b, double x, double y) {
    return tb_relu(__dadd_rn(__dmul_rn(a, x), b)) - y;          // np.maximum(0.0, a*x + b) - ydata
}

// the window statistics the end of a fit needs (tb_lmt_finish_stats): max / min of the bias and the smallest bias under a non-zero signal.
// They depend on the window only, so the pass that forms the first residual norm takes them along and no fit pays a pass for them at its
// end (in the round kernels that pass ran with only the finishing lanes of a warp active -- a third of the instructions of a round).
struct tb_win_stats {
    double xmax, xmin, bmin;
};
__device__ __forceinline__ double tb_lmt_fnorm0_stats(const tb_win_ref W, int m, tb_win_stats &st) {
    tb_enorm_acc n;
    tb_enorm_init(n);
    u32 ymax = 0;
    double xmax = -INFINITY, xmin = INFINITY, bmin = INFINITY;
    TB_FOR_POINTS(W, m, i, {
        ymax = eyi > ymax ? eyi : ymax;
        xmax = fmax(xmax, ex);
        xmin = fmin(xmin, ex);
        if (eyi != 0u) bmin = fmin(bmin, ex);                   // ~isclose(signal, 0)
        tb_enorm_add(n, tb_relu_res(1.0, 1.0, ex, ey), m);
    })
    st.xmax = xmax;
    st.xmin = xmin;
    st.bmin = bmin;
    return ymax > 0 ? tb_enorm_result(n) : -1.0;
}

// residual norm at the start point p0 = (1, 1) and max(signal) of one window; returns -1 when the window holds no signal
__device__ __forceinline__ double tb_lmt_fnorm0(const tb_win_ref W, int m) {
    tb_enorm_acc n;
    tb_enorm_init(n);
    u32 ymax = 0;
    TB_FOR_POINTS(W, m, i, {
        ymax = eyi > ymax ? eyi : ymax;
        tb_enorm_add(n, tb_relu_res(1.0, 1.0, ex, ey), m);
    })
    return ymax > 0 ? tb_enorm_result(n) : -1.0;
}

__device__ __forceinline__ void tb_lmt_init(tb_lmt &S, double fnorm0) {
    S.par0 = 1.0;
    S.par1 = 1.0;
    S.fnorm = fnorm0;
    S.diag0 = S.diag1 = 0.0;
    S.delta = S.lmp = S.xnorm = S.gnorm = 0.0;
    S.r00 = S.r01 = S.r11 = S.qtf0 = S.qtf1 = S.acn0 = S.acn1 = 0.0;
    S.tyy = 0.0;
    S.ipvt0 = 0;
    S.iter = 1;
    S.nfev = 1;
    S.info = 0;
}

// per-point quantities of one outer iteration: residual f, pivot column P and other column Q of the forward-difference
// Jacobian (fdjac2: column j = (fvec(x + h_j e_j) - fvec) / h_j).  (pa, pb) / (qa, qb) are the perturbed parameter
// pairs of the pivot / other column, so the pivoting costs no per-point selects.
struct tb_jac {
    double a, b, pa, pb, ph, phinv, qa, qb, qh, qhinv;
    __device__ __forceinline__ void eval(double x, double y, double &f, double &P, double &Q) const {
        f = tb_relu(__dadd_rn(__dmul_rn(a, x), b)) - y;
        double fp = tb_relu(__dadd_rn(__dmul_rn(pa, x), pb)) - y;
        double fq = tb_relu(__dadd_rn(__dmul_rn(qa, x), qb)) - y;
        P = tb_div_inv(fp - f, ph, phinv);
        Q = tb_div_inv(fq - f, qh, qhinv);
    }
};

// scaled gradient norm, gtol test and the update of the scaling (end of the outer-iteration head, both forms)
__device__ __forceinline__ void tb_lmt_qr_tail(tb_lmt &S, bool swap) {
    const double gtol = 0.0;
    // ---- norm of the scaled gradient
    double gnorm = 0.0;
    if (S.fnorm != 0.0) {
        const double acnp0 = swap ? S.acn1 : S.acn0, acnp1 = swap ? S.acn0 : S.acn1;
        if (acnp0 != 0.0) {
            double s = 0.0;
            s += S.r00 * (S.qtf0 / S.fnorm);
            gnorm = fmax(gnorm, fabs(s / acnp0));
        }
        if (acnp1 != 0.0) {
            double s = 0.0;
            s += S.r01 * (S.qtf0 / S.fnorm);
            s += S.r11 * (S.qtf1 / S.fnorm);
            gnorm = fmax(gnorm, fabs(s / acnp1));
        }
    }
    S.gnorm = gnorm;
    if (gnorm <= gtol) {
        S.info = 4;
        return;
    }
    S.diag0 = fmax(S.diag0, S.acn0);
    S.diag1 = fmax(S.diag1, S.acn1);
}


// fdjac2 + qrfac + Q^T fvec + scaled gradient norm (lmdif outer-loop head).  Sets S.info = 4 when gnorm <= gtol.
__device__ __forceinline__ void tb_lmt_qr(tb_lmt &S, const tb_win_ref W, int m) {
    const double epsmch = 2.220446049250313e-16, factor = 100.0;
    const double eps = sqrt(epsmch);
    tb_jac J;
    J.a = S.par0;
    J.b = S.par1;
    {
        double h = eps * fabs(S.par0);
        if (h == 0.0) h = eps;
        J.ph = h;
        J.pa = S.par0 + h;
        J.pb = S.par1;
        J.phinv = 1.0 / h;
        h = eps * fabs(S.par1);
        if (h == 0.0) h = eps;
        J.qh = h;
        J.qa = S.par0;
        J.qb = S.par1 + h;
        J.qhinv = 1.0 / h;
    }
    S.nfev += 2;
    // the first two rows take part in the Householder bookkeeping (signs, +1 on the diagonal, R's entries)
    const double x0 = W.x[0], y0 = tb_u32_to_double(W.y[0]);
    const double x1 = W.x[W.stride], y1 = tb_u32_to_double(W.y[W.stride]);
    double f0, f1, j00, j10;
    J.eval(x0, y0, f0, j00, j10);
    // ---- pass A: column norms (column 0 = d/da as P, column 1 = d/db as Q)
    tb_enorm_acc n0, n1;
    tb_enorm_init(n0);
    tb_enorm_init(n1);
    TB_FOR_POINTS(W, m, i, {
        double f, j0, j1;
        J.eval(ex, ey, f, j0, j1);
        tb_enorm_add(n0, j0, m);
        tb_enorm_add(n1, j1, m);
    })
    S.acn0 = tb_enorm_result(n0);
    S.acn1 = tb_enorm_result(n1);
    const bool swap = S.acn1 > S.acn0;            // column pivoting
    S.ipvt0 = swap ? 1 : 0;
    if (swap) {
        double t;
        t = J.pa; J.pa = J.qa; J.qa = t;
        t = J.pb; J.pb = J.qb; J.qb = t;
        t = J.ph; J.ph = J.qh; J.qh = t;
        t = J.phinv; J.phinv = J.qhinv; J.qhinv = t;
    }
    double ajnorm = swap ? S.acn1 : S.acn0;
    const bool have0 = ajnorm != 0.0;
    const double P0 = swap ? j10 : j00, Q0 = swap ? j00 : j10;
    double a00 = P0;
    if (have0 && a00 < 0.0) ajnorm = -ajnorm;
    double ainv = 0.0, temp = 0.0, tq0 = 0.0;
    // ---- pass B: Householder vector of the pivot column, its products with the other column and with fvec
    if (have0) {
        ainv = 1.0 / ajnorm;
        a00 = tb_div_inv(P0, ajnorm, ainv) + 1.0;
        double s = 0.0, sf = 0.0, inc = 1.0;
        TB_FOR_POINTS(W, m, i, {
            double f, P, Q;
            J.eval(ex, ey, f, P, Q);
            double c0 = tb_div_inv(P, ajnorm, ainv) + inc;       // + 1 on the first row only (adding 0.0 elsewhere changes no sum)
            inc = 0.0;
            s += __dmul_rn(c0, Q);
            sf += __dmul_rn(c0, f);
        })
        temp = s / a00;
        if (a00 != 0.0) tq0 = -sf / a00;
    }
    S.r00 = -ajnorm;
    const bool app0 = have0 && a00 != 0.0;
    // rows 0 and 1 of the updated second column and of Q0^T fvec
    double P1, Q1;
    J.eval(x1, y1, f1, P1, Q1);
    double c10 = Q0, c11 = Q1, w0 = f0, w1 = f1;
    if (have0) {
        const double c01 = tb_div_inv(P1, ajnorm, ainv);
        c10 = Q0 - __dmul_rn(temp, a00);
        c11 = Q1 - __dmul_rn(temp, c01);
        if (app0) {
            w0 = f0 + __dmul_rn(a00, tq0);
            w1 = f1 + __dmul_rn(c01, tq0);
        }
    }
    S.r01 = c10;
    // ---- pass C: norm of the remaining column below the first row
    tb_enorm_acc nc;
    tb_enorm_init(nc);
    tb_enorm_add(nc, c11, m - 1);
    TB_FOR_POINTS_FROM(W, m, 2, i, {
        double f, P, Q;
        J.eval(ex, ey, f, P, Q);
        double c1 = Q;
        if (have0) c1 = Q - __dmul_rn(temp, tb_div_inv(P, ajnorm, ainv));
        tb_enorm_add(nc, c1, m - 1);
    })
    double ajnorm1 = tb_enorm_result(nc);
    const bool have1 = ajnorm1 != 0.0;
    double ainv1 = 0.0;
    if (have1) {
        if (c11 < 0.0) ajnorm1 = -ajnorm1;
        ainv1 = 1.0 / ajnorm1;
    }
    S.r11 = -ajnorm1;
    // ---- first iteration: scaling and trust-region radius
    if (S.iter == 1) {
        S.diag0 = S.acn0 == 0.0 ? 1.0 : S.acn0;
        S.diag1 = S.acn1 == 0.0 ? 1.0 : S.acn1;
        S.xnorm = tb_enorm2(S.diag0 * S.par0, S.diag1 * S.par1);
        S.delta = factor * S.xnorm;
        if (S.delta == 0.0) S.delta = factor;
    }
    // ---- pass D: qtf = first two components of Q^T fvec
    {
        double a11 = c11;
        if (have1) a11 = tb_div_inv(c11, ajnorm1, ainv1) + 1.0;
        double s1 = 0.0;
        s1 += __dmul_rn(a11, w1);
        TB_FOR_POINTS_FROM(W, m, 2, i, {
            double f, P, Q;
            J.eval(ex, ey, f, P, Q);
            double c1 = Q, w = f;
            if (have0) {
                const double c0 = tb_div_inv(P, ajnorm, ainv);
                c1 = Q - __dmul_rn(temp, c0);
                if (app0) w = f + __dmul_rn(c0, tq0);
            }
            if (have1) c1 = tb_div_inv(c1, ajnorm1, ainv1);
            s1 += __dmul_rn(c1, w);
        })
        S.qtf0 = w0;
        if (a11 != 0.0) {
            double t1 = -s1 / a11;
            w1 += __dmul_rn(a11, t1);
        }
        S.qtf1 = w1;
    }
    tb_lmt_qr_tail(S, swap);
}

// ---- fast form of the outer-iteration head (lm_exact_order = 0) ----------------------------------------------------
// Same LM iteration, same decisions, but R and Q^T fvec come from the five inner products P.P, P.Q, Q.Q, P.f, Q.f of the
// forward-difference Jacobian columns (ONE pass over the window instead of four): with ajnorm = sign(P_0) |P|,
//     r00 = -ajnorm,  r01 = -P.Q / ajnorm,  |r11|^2 = Q.Q - r01^2,  qtf0 = -P.f / ajnorm,  qtf1 = -(Q.f - r01 qtf0) / ajnorm1
// (Householder algebra of qrfac written out for two columns; signs from rows 0 and 1 as MINPACK picks them).  Not bit
// faithful to scipy: R carries a relative error ~ cond(J)^2 * 1e-16 instead of cond(J) * 1e-16, which the forward-difference
// iteration amplifies to ~1e-8 in the fitted parameters -- the same scatter a 1e-13 change of the bias input causes in the
// reference itself, and what the DWM-mode bias input carries anyway.  Structural rank deficiency (fewer than two rows with a
// non-zero Jacobian row) yields r11 = 0 exactly, as in MINPACK.
struct tb_jac_fast {
    double a, b, ap, bp, h0inv, h1inv;
    // residuals exactly as in the exact form (unfused, as numpy evaluates the model)
    __device__ __forceinline__ void eval(double x, double y, double &f, double &j0, double &j1) const {
        const double t = __dmul_rn(a, x);
        f = tb_relu(__dadd_rn(t, b)) - y;
        const double fa = tb_relu(__dadd_rn(__dmul_rn(ap, x), b)) - y;
        const double fb = tb_relu(__dadd_rn(t, bp)) - y;
        j0 = (fa - f) * h0inv;
        j1 = (fb - f) * h1inv;
    }
};

// inner products of the forward-difference Jacobian columns and the residual at one parameter point, plus rows 0 and 1
struct tb_gram {
    double g00, g01, g11, gf0, gf1;
    double f0, a0, b0, f1, a1, b1;                      // rows 0 and 1: residual, d/da, d/db
};
__device__ __forceinline__ tb_jac_fast tb_jac_fast_at(double par0, double par1) {
    const double epsmch = 2.220446049250313e-16;
    const double eps = sqrt(epsmch);
    tb_jac_fast J;
    J.a = par0;
    J.b = par1;
    double h = eps * fabs(par0);
    if (h == 0.0) h = eps;
    J.ap = par0 + h;
    J.h0inv = 1.0 / h;
    h = eps * fabs(par1);
    if (h == 0.0) h = eps;
    J.bp = par1 + h;
    J.h1inv = 1.0 / h;
    return J;
}
__device__ __forceinline__ void tb_lmt_qr_from_gram(tb_lmt &S, const tb_win_ref W, int m, const tb_jac_fast &J, const tb_gram &Gm,
                                                    int nrows_known = -1);

__device__ __forceinline__ void tb_lmt_qr_fast(tb_lmt &S, const tb_win_ref W, int m) {
    const tb_jac_fast J = tb_jac_fast_at(S.par0, S.par1);
    tb_gram Gm;
    J.eval(W.x[0], tb_u32_to_double(W.y[0]), Gm.f0, Gm.a0, Gm.b0);
    J.eval(W.x[W.stride], tb_u32_to_double(W.y[W.stride]), Gm.f1, Gm.a1, Gm.b1);
    double g00 = 0.0, g01 = 0.0, g11 = 0.0, gf0 = 0.0, gf1 = 0.0;
    TB_FOR_POINTS_ANY_ORDER(W, m, {
        double f, j0, j1;
        J.eval(ex, ey, f, j0, j1);
        g00 += j0 * j0;
        g01 += j0 * j1;
        g11 += j1 * j1;
        gf0 += j0 * f;
        gf1 += j1 * f;
    })
    Gm.g00 = g00; Gm.g01 = g01; Gm.g11 = g11; Gm.gf0 = gf0; Gm.gf1 = gf1;
    tb_lmt_qr_from_gram(S, W, m, J, Gm);
}

// One-pass form, start of a fit: the pass that forms the residual norm at p0 = (1, 1) and the window statistics also forms the inner
// products of the first Jacobian (the start point is the same for every window), so the first outer-iteration head costs no pass of its
// own.  Returns -1 for a window without signal (no fit), else the state after the head (S.info may already be non-zero).
__device__ __forceinline__ double tb_lmt_start_fast(tb_lmt &S, const tb_win_ref W, int m, tb_win_stats &st) {
    const tb_jac_fast J = tb_jac_fast_at(1.0, 1.0);
    tb_gram Gm;
    J.eval(W.x[0], tb_u32_to_double(W.y[0]), Gm.f0, Gm.a0, Gm.b0);
    J.eval(W.x[W.stride], tb_u32_to_double(W.y[W.stride]), Gm.f1, Gm.a1, Gm.b1);
    u32 ymax = 0;
    double xmax = -INFINITY, xmin = INFINITY, bmin = INFINITY;
    double ss = 0.0, g00 = 0.0, g01 = 0.0, g11 = 0.0, gf0 = 0.0, gf1 = 0.0;
    TB_FOR_POINTS_ANY_ORDER(W, m, {
        ymax = eyi > ymax ? eyi : ymax;
        xmax = fmax(xmax, ex);
        xmin = fmin(xmin, ex);
        if (eyi != 0u) bmin = fmin(bmin, ex);                   // ~isclose(signal, 0)
        double f, j0, j1;
        J.eval(ex, ey, f, j0, j1);
        ss += f * f;
        g00 += j0 * j0;
        g01 += j0 * j1;
        g11 += j1 * j1;
        gf0 += j0 * f;
        gf1 += j1 * f;
    })
    st.xmax = xmax;
    st.xmin = xmin;
    st.bmin = bmin;
    if (ymax == 0u) return -1.0;
    Gm.g00 = g00; Gm.g01 = g01; Gm.g11 = g11; Gm.gf0 = gf0; Gm.gf1 = gf1;
    const double f0 = sqrt(ss);
    tb_lmt_init(S, f0);
    tb_lmt_qr_from_gram(S, W, m, J, Gm);
    return f0;
}

// R, Q^T fvec, column norms and the outer-iteration bookkeeping from the inner products (no pass over the window unless the columns look
// dependent); counts the two function evaluations of the forward differences
__device__ __forceinline__ void tb_lmt_qr_from_gram(tb_lmt &S, const tb_win_ref W, int m, const tb_jac_fast &J, const tb_gram &Gm,
                                                    int nrows_known) {
    const double factor = 100.0;
    S.nfev += 2;
    const double g00 = Gm.g00, g01 = Gm.g01, g11 = Gm.g11, gf0 = Gm.gf0, gf1 = Gm.gf1;
    const double f0 = Gm.f0, a0 = Gm.a0, b0 = Gm.b0, f1 = Gm.f1, a1 = Gm.a1, b1 = Gm.b1;
    S.acn0 = sqrt(g00);
    S.acn1 = sqrt(g11);
    const bool swap = S.acn1 > S.acn0;
    S.ipvt0 = swap ? 1 : 0;
    const double gPP = swap ? g11 : g00, gQQ = swap ? g00 : g11, gPf = swap ? gf1 : gf0, gQf = swap ? gf0 : gf1;
    const double P0 = swap ? b0 : a0, Q0 = swap ? a0 : b0, P1 = swap ? b1 : a1, Q1 = swap ? a1 : b1;
    double ajnorm = swap ? S.acn1 : S.acn0;
    (void)gPP;
    const bool have0 = ajnorm != 0.0;
    double c11, rest2, dot2, w1;
    if (have0) {
        if (P0 < 0.0) ajnorm = -ajnorm;
        const double ainv = 1.0 / ajnorm;
        S.r01 = -g01 * ainv;
        S.qtf0 = -gPf * ainv;
        const double a00 = P0 * ainv + 1.0;
        const double temp = (g01 * ainv + Q0) / a00;     // (v . Q) / v_0
        const double c01 = P1 * ainv;
        c11 = Q1 - temp * c01;
        const double tq0 = -(gPf * ainv + f0) / a00;
        w1 = f1 + c01 * tq0;
        rest2 = gQQ - S.r01 * S.r01;
        dot2 = gQf - S.r01 * S.qtf0;
    } else {
        S.r01 = Q0;
        S.qtf0 = f0;
        c11 = Q1;
        w1 = f1;
        rest2 = gQQ - Q0 * Q0;
        dot2 = gQf - Q0 * f0;
    }
    S.r00 = -ajnorm;
    // structural rank deficiency (fewer than two rows with a non-zero Jacobian row) must give r11 = 0 exactly, as in MINPACK, where
    // the inner products leave rounding noise: count the rows, but only when the columns look (nearly) dependent
    if (rest2 > 0.0 && !(rest2 > 1e-12 * gQQ)) {
        int nrows = nrows_known;
        if (nrows < 0) {
            nrows = 0;
            TB_FOR_POINTS_ANY_ORDER(W, m, {
                double f, j0, j1;
                J.eval(ex, ey, f, j0, j1);
                nrows += (j0 != 0.0 || j1 != 0.0) ? 1 : 0;
            })
        }
        if (nrows < 2) rest2 = 0.0;
    }
    double ajnorm1 = rest2 > 0.0 ? sqrt(rest2) : 0.0;
    if (ajnorm1 != 0.0) {
        if (c11 < 0.0) ajnorm1 = -ajnorm1;
        S.qtf1 = -dot2 / ajnorm1;
    } else {
        S.qtf1 = w1;
    }
    S.r11 = -ajnorm1;
    if (S.iter == 1) {
        S.diag0 = S.acn0 == 0.0 ? 1.0 : S.acn0;
        S.diag1 = S.acn1 == 0.0 ? 1.0 : S.acn1;
        S.xnorm = tb_enorm2(S.diag0 * S.par0, S.diag1 * S.par1);
        S.delta = factor * S.xnorm;
        if (S.delta == 0.0) S.delta = factor;
    }
    tb_lmt_qr_tail(S, swap);
}

// ---- moment form of the one-pass fits (lm_exact_order = 0, the DWM default) ------------------------------------------------------------
// The model max(0, a x + b) is piecewise linear: on the ACTIVE set A(a, b) = { i : a x_i + b > 0 } the residual is a x_i + b - y_i with the
// Jacobian row (x_i, 1); elsewhere the residual is -y_i and the row is zero.  So everything an LM iteration needs at a parameter point is a
// function of five moments of the active set,  n = |A|, Sx, Sy, Sxx, Sxy,  and of the window constant Tyy = sum of y^2 over all points:
//     |f|^2 = Tyy + a^2 Sxx + 2 a b Sx + b^2 n - 2 a Sxy - 2 b Sy         J'J = [[Sxx, Sx], [Sx, n]]         J'f = (a Sxx + b Sx - Sxy, a Sx + b n - Sy)
// One pass costs a multiply-add for the activity test and five masked accumulations per point (~20 instructions) instead of three model
// evaluations, two difference quotients and six products (~70).  The Jacobian is the exact derivative where scipy takes forward
// differences (h ~ 1.5e-8 |p|): the two agree to the rounding of the difference quotient (~1e-8 relative) except for points within h of
// the kink, and the iteration -- lmpar, trust region, termination tests, the counting of function evaluations (2 per Jacobian, as
// fdjac2) -- is MINPACK's.  Same fit-vs-fallback decisions; parameters as close to scipy's as the forward-difference one-pass form was
// (bench.py `parity`: tracks within the stated tolerance on every window the reference itself keeps stable).
struct tb_moments {
    double sx, sy, sxx, sxy;
    int n;
};
__device__ __forceinline__ bool tb_active(double a, double b, double x) {
    return __dadd_rn(__dmul_rn(a, x), b) > 0.0;                  // np.maximum(0.0, a * x + b) is non-zero
}
__device__ __forceinline__ tb_moments tb_moments_at(const tb_win_ref W, int m, double a, double b) {
    double sx = 0.0, sy = 0.0, sxx = 0.0, sxy = 0.0;
    int n = 0;
    TB_FOR_POINTS_ANY_ORDER(W, m, {
        const bool act = tb_active(a, b, ex);
        const double xm = act ? ex : 0.0;
        const double ym = act ? ey : 0.0;
        n += act ? 1 : 0;
        sx += xm;
        sy += ym;
        sxx = fma(xm, ex, sxx);
        sxy = fma(xm, ey, sxy);
    })
    return tb_moments{sx, sy, sxx, sxy, n};
}
__device__ __forceinline__ double tb_moments_ss(const tb_moments &M, double tyy, double a, double b) {
    const double q = fma(a, fma(a, M.sxx, 2.0 * (b * M.sx - M.sxy)), b * fma(b, (double)M.n, -2.0 * M.sy));
    const double ss = tyy + q;
    return ss > 0.0 ? ss : 0.0;
}
// inner products + rows 0 and 1 at (a, b), in the form tb_lmt_qr_from_gram takes
__device__ __forceinline__ tb_gram tb_moments_gram(const tb_moments &M, const tb_win_ref W, double a, double b) {
    tb_gram G;
    G.g00 = M.sxx;
    G.g01 = M.sx;
    G.g11 = (double)M.n;
    G.gf0 = fma(a, M.sxx, b * M.sx) - M.sxy;
    G.gf1 = fma(a, M.sx, b * (double)M.n) - M.sy;
    const double x0 = W.x[0], y0 = tb_u32_to_double(W.y[0]), x1 = W.x[W.stride], y1 = tb_u32_to_double(W.y[W.stride]);
    const bool a0 = tb_active(a, b, x0), a1 = tb_active(a, b, x1);
    G.f0 = (a0 ? __dadd_rn(__dmul_rn(a, x0), b) : 0.0) - y0;
    G.a0 = a0 ? x0 : 0.0;
    G.b0 = a0 ? 1.0 : 0.0;
    G.f1 = (a1 ? __dadd_rn(__dmul_rn(a, x1), b) : 0.0) - y1;
    G.a1 = a1 ? x1 : 0.0;
    G.b1 = a1 ? 1.0 : 0.0;
    return G;
}
// The exact derivative and scipy's forward differences part ways when the active set is (nearly) degenerate -- fewer than two active points,
// or all of them at the same bias value: the difference quotient also sees the points that the step h pushes across the kink, which can
// make scipy's R non-singular (a usable fit) where the exact Jacobian has rank one (the reference's fallback).  Such heads are rare; they
// are recomputed with the forward-difference inner products, so the fit-vs-fallback decision stays scipy's.
__device__ __forceinline__ bool tb_moments_degenerate(const tb_gram &G) {
    return !(G.g00 * G.g11 - G.g01 * G.g01 > 1e-8 * (G.g00 * G.g11));
}
__device__ __noinline__ void tb_gram_fd(const tb_win_ref W, int m, double a, double b, tb_gram &Gm) {
    const tb_jac_fast J = tb_jac_fast_at(a, b);
    J.eval(W.x[0], tb_u32_to_double(W.y[0]), Gm.f0, Gm.a0, Gm.b0);
    J.eval(W.x[W.stride], tb_u32_to_double(W.y[W.stride]), Gm.f1, Gm.a1, Gm.b1);
    double g00 = 0.0, g01 = 0.0, g11 = 0.0, gf0 = 0.0, gf1 = 0.0;
    TB_FOR_POINTS_ANY_ORDER(W, m, {
        double f, j0, j1;
        J.eval(ex, ey, f, j0, j1);
        g00 += j0 * j0;
        g01 += j0 * j1;
        g11 += j1 * j1;
        gf0 += j0 * f;
        gf1 += j1 * f;
    })
    Gm.g00 = g00; Gm.g01 = g01; Gm.g11 = g11; Gm.gf0 = gf0; Gm.gf1 = gf1;
}
// start of a fit: statistics, Tyy and the moments at p0 = (1, 1) in one pass, then the first outer-iteration head.  -1: no signal.
__device__ __forceinline__ double tb_lmt_start_moments(tb_lmt &S, const tb_win_ref W, int m, tb_win_stats &st) {
    u32 ymax = 0;
    double xmax = -INFINITY, xmin = INFINITY, bmin = INFINITY;
    double sx = 0.0, sy = 0.0, sxx = 0.0, sxy = 0.0, tyy = 0.0;
    int n = 0;
    TB_FOR_POINTS_ANY_ORDER(W, m, {
        ymax = eyi > ymax ? eyi : ymax;
        xmax = fmax(xmax, ex);
        xmin = fmin(xmin, ex);
        if (eyi != 0u) bmin = fmin(bmin, ex);                   // ~isclose(signal, 0)
        const bool act = tb_active(1.0, 1.0, ex);
        const double xm = act ? ex : 0.0;
        const double ym = act ? ey : 0.0;
        n += act ? 1 : 0;
        sx += xm;
        sy += ym;
        sxx = fma(xm, ex, sxx);
        sxy = fma(xm, ey, sxy);
        tyy = fma(ey, ey, tyy);
    })
    st.xmax = xmax;
    st.xmin = xmin;
    st.bmin = bmin;
    if (ymax == 0u) return -1.0;
    const tb_moments M{sx, sy, sxx, sxy, n};
    const double f0 = sqrt(tb_moments_ss(M, tyy, 1.0, 1.0));
    tb_lmt_init(S, f0);
    S.tyy = tyy;
    tb_gram Gm = tb_moments_gram(M, W, 1.0, 1.0);
    int nrows = n;
    if (tb_moments_degenerate(Gm)) {
        tb_gram_fd(W, m, 1.0, 1.0, Gm);
        nrows = -1;
    }
    tb_lmt_qr_from_gram(S, W, m, tb_jac_fast_at(1.0, 1.0), Gm, nrows);
    return f0;
}
// outer-iteration head at the current parameters (finish kernel: a step was accepted in an earlier launch)
__device__ __forceinline__ void tb_lmt_qr_moments(tb_lmt &S, const tb_win_ref W, int m) {
    const tb_moments M = tb_moments_at(W, m, S.par0, S.par1);
    tb_gram Gm = tb_moments_gram(M, W, S.par0, S.par1);
    int nrows = M.n;
    if (tb_moments_degenerate(Gm)) {
        tb_gram_fd(W, m, S.par0, S.par1, Gm);
        nrows = -1;
    }
    tb_lmt_qr_from_gram(S, W, m, tb_jac_fast_at(S.par0, S.par1), Gm, nrows);
}

// one inner iteration of lmdif: lmpar step, trial residual, trust-region update, termination tests.
// returns true when the step was accepted (the next round starts with tb_lmt_qr); S.info != 0 ends the fit.
// FUSE (one-pass form only): the pass that forms the trial residual norm also forms the inner products of the Jacobian columns AT THE
// TRIAL POINT, and when the step is accepted the head of the next outer iteration (tb_lmt_qr_from_gram) follows at once -- an accepted
// step then costs one pass over the window instead of two, and the next round needs no Jacobian pass.  Same arithmetic in the same
// order as the separate passes: the fits are bit-identical.  A rejected step wastes the two extra model evaluations per point.
template <bool FAST, bool FUSE = false, bool MOMENTS = false>
__device__ __forceinline__ bool tb_lmt_trial(tb_lmt &S, const tb_win_ref W, int m) {
    const double epsmch = 2.220446049250313e-16;
    const double ftol = 1.49012e-8, xtol = 1.49012e-8;
    const double p1 = 0.1, p5 = 0.5, p25 = 0.25, p75 = 0.75, p0001 = 1e-4;
    const int maxfev = 600;
    double r[2][2] = {{S.r00, S.r01}, {0.0, S.r11}};
    const int ipvt[2] = {S.ipvt0, 1 - S.ipvt0};
    const double diag[2] = {S.diag0, S.diag1};
    const double qtf[2] = {S.qtf0, S.qtf1};
    double wa1[2], wa2[2], wa3[2];
    tb_lmpar2(r, ipvt, diag, qtf, S.delta, S.lmp, wa1, wa2);
    const double par[2] = {S.par0, S.par1};
    for (int j = 0; j < 2; j++) {
        wa1[j] = -wa1[j];
        wa2[j] = par[j] + wa1[j];
        wa3[j] = diag[j] * wa1[j];
    }
    double pnorm = tb_enorm2(wa3[0], wa3[1]);
    if (S.iter == 1) S.delta = fmin(S.delta, pnorm);
    double fnorm1;
    tb_jac_fast J;
    tb_gram Gm;
    int nrows = -1;
    {
        const double ta = wa2[0], tb_ = wa2[1];
        if (MOMENTS) {                               // moment form: always fused (the moments give the residual norm AND the next head)
            const tb_moments M = tb_moments_at(W, m, ta, tb_);
            fnorm1 = sqrt(tb_moments_ss(M, S.tyy, ta, tb_));
            Gm = tb_moments_gram(M, W, ta, tb_);
            J = tb_jac_fast_at(ta, tb_);
            nrows = M.n;
            if (tb_moments_degenerate(Gm)) {
                tb_gram_fd(W, m, ta, tb_, Gm);
                nrows = -1;
            }
        } else if (FAST && FUSE) {
            J = tb_jac_fast_at(ta, tb_);
            J.eval(W.x[0], tb_u32_to_double(W.y[0]), Gm.f0, Gm.a0, Gm.b0);
            J.eval(W.x[W.stride], tb_u32_to_double(W.y[W.stride]), Gm.f1, Gm.a1, Gm.b1);
            double ss = 0.0, g00 = 0.0, g01 = 0.0, g11 = 0.0, gf0 = 0.0, gf1 = 0.0;
            TB_FOR_POINTS_ANY_ORDER(W, m, {
                double f, j0, j1;
                J.eval(ex, ey, f, j0, j1);
                ss += f * f;
                g00 += j0 * j0;
                g01 += j0 * j1;
                g11 += j1 * j1;
                gf0 += j0 * f;
                gf1 += j1 * f;
            })
            Gm.g00 = g00; Gm.g01 = g01; Gm.g11 = g11; Gm.gf0 = gf0; Gm.gf1 = gf1;
            fnorm1 = sqrt(ss);
        } else if (FAST) {
            double ss = 0.0;
            TB_FOR_POINTS_ANY_ORDER(W, m, {
                const double r = tb_relu_res(ta, tb_, ex, ey);
                ss += r * r;
            })
            fnorm1 = sqrt(ss);
        } else {
            tb_enorm_acc n;
            tb_enorm_init(n);
            TB_FOR_POINTS(W, m, i, { tb_enorm_add(n, tb_relu_res(ta, tb_, ex, ey), m); })
            fnorm1 = tb_enorm_result(n);
        }
    }
    S.nfev++;
    double actred = -1.0;
    if (p1 * fnorm1 < S.fnorm) {
        double t = fnorm1 / S.fnorm;
        actred = 1.0 - t * t;
    }
    r[0][0] = S.r00;
    r[0][1] = S.r01;
    r[1][1] = S.r11;
    for (int j = 0; j < 2; j++) {
        wa3[j] = 0.0;
        double temp = wa1[ipvt[j]];
        for (int i = 0; i <= j; i++) wa3[i] += r[i][j] * temp;
    }
    double temp1 = tb_enorm2(wa3[0], wa3[1]) / S.fnorm;
    double temp2 = sqrt(S.lmp) * pnorm / S.fnorm;
    double prered = temp1 * temp1 + temp2 * temp2 / p5;
    double dirder = -(temp1 * temp1 + temp2 * temp2);
    double ratio = 0.0;
    if (prered != 0.0) ratio = actred / prered;
    if (ratio <= p25) {
        double temp;
        if (actred >= 0.0) temp = p5;
        else temp = p5 * dirder / (dirder + p5 * actred);
        if (p1 * fnorm1 >= S.fnorm || temp < p1) temp = p1;
        S.delta = temp * fmin(S.delta, pnorm / p1);
        S.lmp /= temp;
    } else if (S.lmp == 0.0 || ratio >= p75) {
        S.delta = pnorm / p5;
        S.lmp = p5 * S.lmp;
    }
    const bool accepted = ratio >= p0001;
    if (accepted) {
        S.par0 = wa2[0];
        S.par1 = wa2[1];
        S.xnorm = tb_enorm2(S.diag0 * S.par0, S.diag1 * S.par1);
        S.fnorm = fnorm1;
        S.iter++;
    }
    int info = 0;
    if (fabs(actred) <= ftol && prered <= ftol && p5 * ratio <= 1.0) info = 1;
    if (S.delta <= xtol * S.xnorm) info = 2;
    if (fabs(actred) <= ftol && prered <= ftol && p5 * ratio <= 1.0 && info == 2) info = 3;
    if (info == 0) {
        if (S.nfev >= maxfev) info = 5;
        if (fabs(actred) <= epsmch && prered <= epsmch && p5 * ratio <= 1.0) info = 6;
        if (S.delta <= epsmch * S.xnorm) info = 7;
        if (S.gnorm <= epsmch) info = 8;
    }
    S.info = info;
    if ((MOMENTS || (FAST && FUSE)) && accepted && info == 0) tb_lmt_qr_from_gram(S, W, m, J, Gm, nrows);      // head of the next outer iteration (may set S.info)
    return accepted;
}

// ---- one-pass form as ONE step function (persistent kernel, tb_correct.cu) -------------------------------------------------------
// A step = [where to evaluate] -> ONE pass over the window -> [what the pass means].  first: the start of a fit -- the pass runs at
// p0 = (1, 1), also takes the window statistics, and is followed by tb_lmt_init + the first outer-iteration head (= tb_lmt_start_fast).
// Otherwise: lmpar step -> pass at the trial point (residual norm + inner products of the Jacobian there) -> trust-region update, and
// the head of the next outer iteration when the step was accepted (= tb_lmt_trial<true, true>).  Lanes of a warp that are at different
// stages of different fits share the pass -- the expensive part -- and diverge only in the scalar algebra around it.
// Same arithmetic in the same order as the two functions it unites: bit-identical fits.  Returns false for a window without signal.
__device__ __forceinline__ bool tb_lmt_step_fast(tb_lmt &S, tb_win_stats &st, const tb_win_ref W, int m, bool first) {
    const double epsmch = 2.220446049250313e-16;
    const double ftol = 1.49012e-8, xtol = 1.49012e-8;
    const double p1 = 0.1, p5 = 0.5, p25 = 0.25, p75 = 0.75, p0001 = 1e-4;
    const int maxfev = 600;
    double wa1[2] = {0.0, 0.0}, wa2[2] = {1.0, 1.0}, pnorm = 0.0;
    int ipvt[2] = {0, 1};
    if (!first) {
        double r[2][2] = {{S.r00, S.r01}, {0.0, S.r11}};
        ipvt[0] = S.ipvt0;
        ipvt[1] = 1 - S.ipvt0;
        const double diag[2] = {S.diag0, S.diag1};
        const double qtf[2] = {S.qtf0, S.qtf1};
        double wa3[2];
        tb_lmpar2(r, ipvt, diag, qtf, S.delta, S.lmp, wa1, wa2);
        const double par[2] = {S.par0, S.par1};
        for (int j = 0; j < 2; j++) {
            wa1[j] = -wa1[j];
            wa2[j] = par[j] + wa1[j];
            wa3[j] = diag[j] * wa1[j];
        }
        pnorm = tb_enorm2(wa3[0], wa3[1]);
        if (S.iter == 1) S.delta = fmin(S.delta, pnorm);
    }
    // ---- the pass
    const tb_jac_fast J = tb_jac_fast_at(wa2[0], wa2[1]);
    tb_gram Gm;
    J.eval(W.x[0], tb_u32_to_double(W.y[0]), Gm.f0, Gm.a0, Gm.b0);
    J.eval(W.x[W.stride], tb_u32_to_double(W.y[W.stride]), Gm.f1, Gm.a1, Gm.b1);
    u32 ymax = 0;
    double xmax = -INFINITY, xmin = INFINITY, bmin = INFINITY;
    double ss = 0.0, g00 = 0.0, g01 = 0.0, g11 = 0.0, gf0 = 0.0, gf1 = 0.0;
    TB_FOR_POINTS_ANY_ORDER(W, m, {
        if (first) {
            ymax = eyi > ymax ? eyi : ymax;
            xmax = fmax(xmax, ex);
            xmin = fmin(xmin, ex);
            if (eyi != 0u) bmin = fmin(bmin, ex);               // ~isclose(signal, 0)
        }
        double f, j0, j1;
        J.eval(ex, ey, f, j0, j1);
        ss += f * f;
        g00 += j0 * j0;
        g01 += j0 * j1;
        g11 += j1 * j1;
        gf0 += j0 * f;
        gf1 += j1 * f;
    })
    Gm.g00 = g00; Gm.g01 = g01; Gm.g11 = g11; Gm.gf0 = gf0; Gm.gf1 = gf1;
    const double fnorm1 = sqrt(ss);
    if (first) {
        st.xmax = xmax;
        st.xmin = xmin;
        st.bmin = bmin;
        if (ymax == 0u) return false;
        tb_lmt_init(S, fnorm1);
        tb_lmt_qr_from_gram(S, W, m, J, Gm);
        return true;
    }
    // ---- what the trial means (lmdif inner loop)
    S.nfev++;
    double actred = -1.0;
    if (p1 * fnorm1 < S.fnorm) {
        double t = fnorm1 / S.fnorm;
        actred = 1.0 - t * t;
    }
    double wa3[2];
    {
        const double r[2][2] = {{S.r00, S.r01}, {0.0, S.r11}};
        for (int j = 0; j < 2; j++) {
            wa3[j] = 0.0;
            double temp = wa1[ipvt[j]];
            for (int i = 0; i <= j; i++) wa3[i] += r[i][j] * temp;
        }
    }
    double temp1 = tb_enorm2(wa3[0], wa3[1]) / S.fnorm;
    double temp2 = sqrt(S.lmp) * pnorm / S.fnorm;
    double prered = temp1 * temp1 + temp2 * temp2 / p5;
    double dirder = -(temp1 * temp1 + temp2 * temp2);
    double ratio = 0.0;
    if (prered != 0.0) ratio = actred / prered;
    if (ratio <= p25) {
        double temp;
        if (actred >= 0.0) temp = p5;
        else temp = p5 * dirder / (dirder + p5 * actred);
        if (p1 * fnorm1 >= S.fnorm || temp < p1) temp = p1;
        S.delta = temp * fmin(S.delta, pnorm / p1);
        S.lmp /= temp;
    } else if (S.lmp == 0.0 || ratio >= p75) {
        S.delta = pnorm / p5;
        S.lmp = p5 * S.lmp;
    }
    const bool accepted = ratio >= p0001;
    if (accepted) {
        S.par0 = wa2[0];
        S.par1 = wa2[1];
        S.xnorm = tb_enorm2(S.diag0 * S.par0, S.diag1 * S.par1);
        S.fnorm = fnorm1;
        S.iter++;
    }
    int info = 0;
    if (fabs(actred) <= ftol && prered <= ftol && p5 * ratio <= 1.0) info = 1;
    if (S.delta <= xtol * S.xnorm) info = 2;
    if (fabs(actred) <= ftol && prered <= ftol && p5 * ratio <= 1.0 && info == 2) info = 3;
    if (info == 0) {
        if (S.nfev >= maxfev) info = 5;
        if (fabs(actred) <= epsmch && prered <= epsmch && p5 * ratio <= 1.0) info = 6;
        if (S.delta <= epsmch * S.xnorm) info = 7;
        if (S.gnorm <= epsmch) info = 8;
    }
    S.info = info;
    if (accepted && info == 0) tb_lmt_qr_from_gram(S, W, m, J, Gm);      // head of the next outer iteration (may set S.info)
    return true;
}

// scipy's post-conditions (ier in {1,2,3,4}; inv(R) exists and the covariance holds no NaN) and the window record:
// mode 0: usable fit (a, b), mode 1: the reference's fallback (atacorrect_functions.py:296-299); pmax = max(prediction).
// The prediction max(0, a*x + b) is monotonic in x (rounding is monotonic), so its maximum sits at max(x) or min(x).
__device__ __forceinline__ void tb_lmt_finish_stats(const tb_lmt &S, const tb_win_stats &st, double &mode, double &pa, double &pb, double &pmax) {
    int usable = (S.info >= 1 && S.info <= 4) ? 1 : 0;
    if (usable) {
        if (S.r00 == 0.0 || S.r11 == 0.0) usable = 0;
        else {
            double i00 = 1.0 / S.r00, i11 = 1.0 / S.r11;
            double i01 = -i11 * (i00 * S.r01);
            double c00 = i00 * i00 + i01 * i01, c01 = i01 * i11, c11 = i11 * i11;
            if (c00 != c00 || c01 != c01 || c11 != c11) usable = 0;
        }
    }
    const double xmax = st.xmax, xmin = st.xmin, bmin = st.bmin;
    if (usable) {
        const double a = S.par0, b = S.par1;
        mode = 0.0;
        pa = a;
        pb = b;
        double t = __dadd_rn(__dmul_rn(a, a < 0.0 ? xmin : xmax), b);
        pmax = t > 0.0 ? t : 0.0;
    } else {
        mode = 1.0;
        pa = 0.0;
        pb = bmin;
        double t = xmax - bmin;                                 // max_i max(0, x_i - bmin)
        pmax = t < 0.0 ? 0.0 : t;
    }
}
__device__ __forceinline__ void tb_lmt_finish(const tb_lmt &S, const tb_win_ref W, int m, double &mode, double &pa, double &pb,
                                              double &pmax) {
    tb_win_stats st{-INFINITY, INFINITY, INFINITY};
    TB_FOR_POINTS(W, m, i, {
        st.xmax = fmax(st.xmax, ex);
        st.xmin = fmin(st.xmin, ex);
        if (eyi != 0u) st.bmin = fmin(st.bmin, ex);             // ~isclose(signal, 0)
        (void)ey;
    })
    tb_lmt_finish_stats(S, st, mode, pa, pb, pmax);
}

// complete fit of one window by one thread (test harness entry; the kernel interleaves the phases across lanes)
template <bool FAST, bool FUSE = false>
__device__ __forceinline__ void tb_lm_fit_thread(const tb_win_ref W, int m, tb_lm_result &res) {
    tb_lmt S;
    double f0 = tb_lmt_fnorm0(W, m);
    tb_lmt_init(S, f0 < 0.0 ? 0.0 : f0);
    bool need_qr = true;
    for (;;) {
        if (need_qr) {
            if (FAST) tb_lmt_qr_fast(S, W, m);
            else tb_lmt_qr(S, W, m);
            if (S.info) break;
        }
        const bool accepted = tb_lmt_trial<FAST, FAST && FUSE>(S, W, m);
        if (S.info) break;
        need_qr = accepted && !(FAST && FUSE);          // fused: the head of the next outer iteration ran inside the trial
    }
    double mode, pa, pb, pmax;
    tb_lmt_finish(S, W, m, mode, pa, pb, pmax);
    res.a = S.par0;
    res.b = S.par1;
    res.info = S.info;
    res.nfev = S.nfev;
    res.usable = mode == 0.0 ? 1 : 0;
}
