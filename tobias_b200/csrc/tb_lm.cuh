// tb_lm.cuh -- K3b core: warp-cooperative Levenberg-Marquardt fit of  signal ~ max(0, a*bias + b)  on one window.
//
// This replaces scipy.optimize.curve_fit(relu, bias_w, signal_w) as called by the reference at
// tobias/tools/atacorrect_functions.py:292 (method 'lm' -> scipy.optimize.leastsq -> MINPACK lmdif; scipy is
// unpinned by the reference, 1.18.1 in this image).  The published MINPACK algorithm (More, Garbow, Hillstrom:
// lmdif / fdjac2 / qrfac / qrsolv / lmpar / enorm, Netlib) is restated for n = 2 parameters with scipy's settings:
// p0 = (1, 1), ftol = xtol = 1.49012e-8, gtol = 0, maxfev = 200*(n+1) = 600, epsfcn = machine eps, factor = 100,
// mode = 1 (internal scaling)  (scipy/optimize/_minpack_py.py:292-294,433-438,919-921).
//
// One warp fits one window: lane l holds points l, l+32, ...; sums over the m points are warp reductions
// (xor butterflies, identical result in every lane), the 2x2 trust-region algebra runs redundantly in all lanes.
// The iteration follows MINPACK decision by decision (same tests, same constants); only the association order
// of the m-term sums differs from the sequential Fortran/C loops (~1 ulp), which leaves the iteration path
// unchanged except when a test lands within rounding of its threshold.
//
// The model is evaluated as the reference's numpy expression does: t = a*x (rounded), t + b (rounded), max(0, .) --
// no fused multiply-add (this translation unit is compiled with -fmad=false; the intrinsics make it explicit).
#pragma once
#include "tb_common.cuh"

#define TB_LM_FULL 0xffffffffu

__device__ __forceinline__ double tb_warp_sum(double v) {
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(TB_LM_FULL, v, d);
    return v;
}
__device__ __forceinline__ double tb_warp_max(double v) {
    for (int d = 16; d > 0; d >>= 1) {
        double o = __shfl_xor_sync(TB_LM_FULL, v, d);
        v = o > v ? o : v;
    }
    return v;
}
__device__ __forceinline__ double tb_warp_min(double v) {
    for (int d = 16; d > 0; d >>= 1) {
        double o = __shfl_xor_sync(TB_LM_FULL, v, d);
        v = o < v ? o : v;
    }
    return v;
}

#ifdef TB_FIRST_GEN_FITS      // warp-per-window kernels of the first generation: test / comparison builds only (tests/emu, variants/)
// MINPACK enorm (three accumulators: large / intermediate / small components), sequential over v[first..m).
__device__ __forceinline__ double tb_enorm_seq(const double *v, int first, int m) {
    const double rdwarf = 3.834e-20, rgiant = 1.304e19;
    double s1 = 0, s2 = 0, s3 = 0, x1max = 0, x3max = 0;
    const double agiant = rgiant / (double)(m - first);
    for (int i = first; i < m; i++) {
        double xabs = fabs(v[i]);
        if (xabs > rdwarf && xabs < agiant) {
            s2 += xabs * xabs;
        } else if (xabs <= rdwarf) {
            if (xabs > x3max) {
                double t = x3max / xabs;
                s3 = 1.0 + s3 * (t * t);
                x3max = xabs;
            } else if (xabs != 0.0) {
                double t = xabs / x3max;
                s3 += t * t;
            }
        } else {
            if (xabs > x1max) {
                double t = x1max / xabs;
                s1 = 1.0 + s1 * (t * t);
                x1max = xabs;
            } else {
                double t = xabs / x1max;
                s1 += t * t;
            }
        }
    }
    if (s1 != 0.0) return x1max * sqrt(s1 + (s2 / x1max) / x1max);
    if (s2 != 0.0) {
        if (s2 >= x3max) return sqrt(s2 * (1.0 + (x3max / s2) * (x3max * s3)));
        return sqrt(x3max * ((s2 / x3max) + (x3max * s3)));
    }
    return x3max * sqrt(s3);
}

// stage this lane's elements of an m-vector into the warp's scratch row (index order), visible to the whole warp
template <int PPL>
__device__ __forceinline__ void tb_stage(const double (&v)[PPL], int lane, double *wbuf) {
    __syncwarp();
#pragma unroll
    for (int q = 0; q < PPL; q++) wbuf[lane + 32 * q] = v[q];
    __syncwarp();
}

// Euclidean norm of the m-vector distributed over the warp; elements with index < first are excluded.
// SEQ: exact MINPACK order (every lane walks the staged vector); else warp-tree reduction.
template <int PPL, bool SEQ>
__device__ __forceinline__ double tb_enorm_warp(const double (&v)[PPL], int m, int lane, int first, double *wbuf) {
    if (SEQ) {
        tb_stage<PPL>(v, lane, wbuf);
        return tb_enorm_seq(wbuf, first, m);
    }
    double amax = 0.0;
#pragma unroll
    for (int q = 0; q < PPL; q++) {
        int i = lane + 32 * q;
        if (i < m && i >= first) amax = fmax(amax, fabs(v[q]));
    }
    amax = tb_warp_max(amax);
    if (amax == 0.0) return 0.0;
    double s = 0.0;
    if (amax > 1e-140 && amax < 1e140) {
#pragma unroll
        for (int q = 0; q < PPL; q++) {
            int i = lane + 32 * q;
            if (i < m && i >= first) s += v[q] * v[q];
        }
        return sqrt(tb_warp_sum(s));
    }
#pragma unroll
    for (int q = 0; q < PPL; q++) {
        int i = lane + 32 * q;
        if (i < m && i >= first) {
            double t = v[q] / amax;
            s += t * t;
        }
    }
    return amax * sqrt(tb_warp_sum(s));
}

// sum_{i = first}^{m-1} a[i] * b[i]   (each product rounded, then added: no fused multiply-add)
template <int PPL, bool SEQ>
__device__ __forceinline__ double tb_dot_warp(const double (&a)[PPL], const double (&b)[PPL], int m, int lane, int first,
                                              double *wbuf) {
    if (SEQ) {
        double prod[PPL];
#pragma unroll
        for (int q = 0; q < PPL; q++) prod[q] = __dmul_rn(a[q], b[q]);
        tb_stage<PPL>(prod, lane, wbuf);
        double s = 0.0;
        for (int i = first; i < m; i++) s += wbuf[i];
        return s;
    }
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < PPL; q++) {
        int i = lane + 32 * q;
        if (i < m && i >= first) s += a[q] * b[q];
    }
    return tb_warp_sum(s);
}

#endif  // TB_FIRST_GEN_FITS

// MINPACK enorm for the 2-vectors of the trust-region algebra (three-accumulator form).
__device__ __noinline__ double tb_enorm2(double x0, double x1) {
    const double rdwarf = 3.834e-20, rgiant = 1.304e19;
    double s1 = 0, s2 = 0, s3 = 0, x1max = 0, x3max = 0;
    const double agiant = rgiant / 2.0;
    double xs[2] = {x0, x1};
    for (int i = 0; i < 2; i++) {
        double xabs = fabs(xs[i]);
        if (xabs > rdwarf && xabs < agiant) {
            s2 += xabs * xabs;
        } else if (xabs <= rdwarf) {
            if (xabs > x3max) {
                double t = x3max / xabs;
                s3 = 1.0 + s3 * (t * t);
                x3max = xabs;
            } else if (xabs != 0.0) {
                double t = xabs / x3max;
                s3 += t * t;
            }
        } else {
            if (xabs > x1max) {
                double t = x1max / xabs;
                s1 = 1.0 + s1 * (t * t);
                x1max = xabs;
            } else {
                double t = xabs / x1max;
                s1 += t * t;
            }
        }
    }
    if (s1 != 0.0) return x1max * sqrt(s1 + (s2 / x1max) / x1max);
    if (s2 != 0.0) {
        if (s2 >= x3max) return sqrt(s2 * (1.0 + (x3max / s2) * (x3max * s3)));
        return sqrt(x3max * ((s2 / x3max) + (x3max * s3)));
    }
    return x3max * sqrt(s3);
}

// qrsolv for n = 2.  r: upper triangle = R (full diagonal), strict lower triangle is scratch.
__device__ __forceinline__ void tb_qrsolv2(double (&r)[2][2], const int (&ipvt)[2], const double (&diag)[2],
                                           const double (&qtb)[2], double (&x)[2], double (&sdiag)[2]) {
    double wa[2];
    for (int j = 0; j < 2; j++) {
        for (int i = j; i < 2; i++) r[i][j] = r[j][i];
        x[j] = r[j][j];
        wa[j] = qtb[j];
    }
    for (int j = 0; j < 2; j++) {
        int l = ipvt[j];
        if (diag[l] != 0.0) {
            for (int k = j; k < 2; k++) sdiag[k] = 0.0;
            sdiag[j] = diag[l];
            double qtbpj = 0.0;
            for (int k = j; k < 2; k++) {
                if (sdiag[k] == 0.0) continue;
                double c, s;
                if (fabs(r[k][k]) < fabs(sdiag[k])) {
                    double cotan = r[k][k] / sdiag[k];
                    s = 0.5 / sqrt(0.25 + 0.25 * (cotan * cotan));
                    c = s * cotan;
                } else {
                    double t = sdiag[k] / r[k][k];
                    c = 0.5 / sqrt(0.25 + 0.25 * (t * t));
                    s = c * t;
                }
                r[k][k] = c * r[k][k] + s * sdiag[k];
                double temp = c * wa[k] + s * qtbpj;
                qtbpj = -s * wa[k] + c * qtbpj;
                wa[k] = temp;
                for (int i = k + 1; i < 2; i++) {
                    temp = c * r[i][k] + s * sdiag[i];
                    sdiag[i] = -s * r[i][k] + c * sdiag[i];
                    r[i][k] = temp;
                }
            }
        }
        sdiag[j] = r[j][j];
        r[j][j] = x[j];
    }
    int nsing = 2;
    for (int j = 0; j < 2; j++) {
        if (sdiag[j] == 0.0 && nsing == 2) nsing = j;
        if (nsing < 2) wa[j] = 0.0;
    }
    for (int k = 1; k <= nsing; k++) {
        int j = nsing - k;
        double sum = 0.0;
        for (int i = j + 1; i < nsing; i++) sum += r[i][j] * wa[i];
        wa[j] = (wa[j] - sum) / sdiag[j];
    }
    for (int j = 0; j < 2; j++) x[ipvt[j]] = wa[j];
}

// lmpar for n = 2.
__device__ __noinline__ void tb_lmpar2(double (&r)[2][2], const int (&ipvt)[2], const double (&diag)[2],
                                          const double (&qtb)[2], double delta, double &par, double (&x)[2],
                                          double (&sdiag)[2]) {
    const double dwarf = 2.2250738585072014e-308, p1 = 0.1, p001 = 0.001;
    double wa1[2], wa2[2];
    int nsing = 2;
    for (int j = 0; j < 2; j++) {
        wa1[j] = qtb[j];
        if (r[j][j] == 0.0 && nsing == 2) nsing = j;
        if (nsing < 2) wa1[j] = 0.0;
    }
    for (int k = 1; k <= nsing; k++) {
        int j = nsing - k;
        wa1[j] /= r[j][j];
        double temp = wa1[j];
        for (int i = 0; i < j; i++) wa1[i] -= r[i][j] * temp;
    }
    for (int j = 0; j < 2; j++) x[ipvt[j]] = wa1[j];
    int iter = 0;
    for (int j = 0; j < 2; j++) wa2[j] = diag[j] * x[j];
    double dxnorm = tb_enorm2(wa2[0], wa2[1]);
    double fp = dxnorm - delta;
    if (fp <= p1 * delta) {
        par = 0.0;
        return;
    }
    double parl = 0.0;
    if (nsing >= 2) {
        for (int j = 0; j < 2; j++) {
            int l = ipvt[j];
            wa1[j] = diag[l] * (wa2[l] / dxnorm);
        }
        for (int j = 0; j < 2; j++) {
            double sum = 0.0;
            for (int i = 0; i < j; i++) sum += r[i][j] * wa1[i];
            wa1[j] = (wa1[j] - sum) / r[j][j];
        }
        double temp = tb_enorm2(wa1[0], wa1[1]);
        parl = fp / delta / temp / temp;
    }
    for (int j = 0; j < 2; j++) {
        double sum = 0.0;
        for (int i = 0; i <= j; i++) sum += r[i][j] * qtb[i];
        wa1[j] = sum / diag[ipvt[j]];
    }
    double gnorm = tb_enorm2(wa1[0], wa1[1]);
    double paru = gnorm / delta;
    if (paru == 0.0) paru = dwarf / fmin(delta, p1);
    par = fmax(par, parl);
    par = fmin(par, paru);
    if (par == 0.0) par = gnorm / dxnorm;
    for (;;) {
        iter++;
        if (par == 0.0) par = fmax(dwarf, p001 * paru);
        double temp = sqrt(par);
        for (int j = 0; j < 2; j++) wa1[j] = temp * diag[j];
        tb_qrsolv2(r, ipvt, wa1, qtb, x, sdiag);
        for (int j = 0; j < 2; j++) wa2[j] = diag[j] * x[j];
        dxnorm = tb_enorm2(wa2[0], wa2[1]);
        temp = fp;
        fp = dxnorm - delta;
        if (fabs(fp) <= p1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10) break;
        for (int j = 0; j < 2; j++) {
            int l = ipvt[j];
            wa1[j] = diag[l] * (wa2[l] / dxnorm);
        }
        for (int j = 0; j < 2; j++) {
            wa1[j] /= sdiag[j];
            double t = wa1[j];
            for (int i = j + 1; i < 2; i++) wa1[i] -= r[i][j] * t;
        }
        temp = tb_enorm2(wa1[0], wa1[1]);
        double parc = fp / delta / temp / temp;
        if (fp > 0.0) parl = fmax(parl, par);
        if (fp < 0.0) paru = fmin(paru, par);
        par = fmax(parl, par + parc);
    }
}

struct tb_lm_result {
    double a, b;
    int info;        // MINPACK info (1..8)
    int nfev;
    int usable;      // 1: curve_fit would return popt; 0: it raises (RuntimeError / OptimizeWarning) -> fallback
};

#ifdef TB_FIRST_GEN_FITS
template <int PPL>
__device__ __forceinline__ void tb_relu_residuals(const double (&x)[PPL], const double (&y)[PPL], double a, double b,
                                                  double (&f)[PPL]) {
#ifdef TB_LM_TRACE
    TB_LM_TRACE(a, b);      // debug harness hook (never defined in the library build)
#endif
#pragma unroll
    for (int q = 0; q < PPL; q++) {
        double t = __dadd_rn(__dmul_rn(a, x[q]), b);
        f[q] = (t > 0.0 ? t : 0.0) - y[q];       // np.maximum(0.0, a*x + b) - ydata
    }
}

template <int PPL>
__device__ __forceinline__ double tb_elem(const double (&v)[PPL], int idx) {   // value of element idx, in every lane
    return __shfl_sync(TB_LM_FULL, v[idx >> 5 < PPL ? idx >> 5 : 0], idx & 31);
}

// x, y: this lane's points (index lane + 32 q), m = number of points in the window (2 <= m <= 32 * PPL).
// wbuf: 32*PPL doubles of shared memory private to this warp (only touched when SEQ)
template <int PPL, bool SEQ>
__device__ void tb_lm_fit_warp(const double (&x)[PPL], const double (&y)[PPL], int m, int lane, double *wbuf,
                               tb_lm_result &res) {
    const double epsmch = 2.220446049250313e-16;
    const double ftol = 1.49012e-8, xtol = 1.49012e-8, gtol = 0.0, factor = 100.0;
    const double p1 = 0.1, p5 = 0.5, p25 = 0.25, p75 = 0.75, p0001 = 1e-4;
    const int maxfev = 600;
    double par[2] = {1.0, 1.0};
    double fvec[PPL], wa4[PPL], c0[PPL], c1[PPL];
    double diag[2] = {0, 0}, qtf[2], wa1[2], wa2[2], wa3[2], acn[2];
    double r[2][2] = {{0, 0}, {0, 0}};
    int ipvt[2] = {0, 1};
    int info = 0, nfev = 0, iter = 1;
    double lmp = 0.0, delta = 0.0, xnorm = 0.0, gnorm = 0.0;

    tb_relu_residuals<PPL>(x, y, par[0], par[1], fvec);
    nfev = 1;
    double fnorm = tb_enorm_warp<PPL, SEQ>(fvec, m, lane, 0, wbuf);
    const double eps = sqrt(epsmch);

    for (;;) {
        // ---- fdjac2: forward-difference Jacobian, columns c0 (d/da) and c1 (d/db)
        {
            double temp = par[0], h = eps * fabs(temp);
            if (h == 0.0) h = eps;
            tb_relu_residuals<PPL>(x, y, temp + h, par[1], c0);
            double hinv = 1.0 / h;            // tree mode only: one reciprocal instead of PPL divisions (<= 1 ulp, like the sums)
#pragma unroll
            for (int q = 0; q < PPL; q++) c0[q] = SEQ ? (c0[q] - fvec[q]) / h : (c0[q] - fvec[q]) * hinv;
            temp = par[1];
            h = eps * fabs(temp);
            if (h == 0.0) h = eps;
            tb_relu_residuals<PPL>(x, y, par[0], temp + h, c1);
            hinv = 1.0 / h;
#pragma unroll
            for (int q = 0; q < PPL; q++) c1[q] = SEQ ? (c1[q] - fvec[q]) / h : (c1[q] - fvec[q]) * hinv;
            nfev += 2;
        }
        // ---- qrfac with column pivoting
        acn[0] = tb_enorm_warp<PPL, SEQ>(c0, m, lane, 0, wbuf);
        acn[1] = tb_enorm_warp<PPL, SEQ>(c1, m, lane, 0, wbuf);
        ipvt[0] = 0;
        ipvt[1] = 1;
        double rdiag0 = acn[0];
        if (acn[1] > acn[0]) {
#pragma unroll
            for (int q = 0; q < PPL; q++) {
                double t = c0[q];
                c0[q] = c1[q];
                c1[q] = t;
            }
            ipvt[0] = 1;
            ipvt[1] = 0;
            rdiag0 = acn[1];
        }
        double ajnorm = rdiag0;
        if (ajnorm != 0.0) {
            if (tb_elem<PPL>(c0, 0) < 0.0) ajnorm = -ajnorm;
            const double ainv = 1.0 / ajnorm;
#pragma unroll
            for (int q = 0; q < PPL; q++) c0[q] = SEQ ? c0[q] / ajnorm : c0[q] * ainv;
            if (lane == 0) c0[0] += 1.0;
            double s = tb_dot_warp<PPL, SEQ>(c0, c1, m, lane, 0, wbuf);
            double temp = s / tb_elem<PPL>(c0, 0);
#pragma unroll
            for (int q = 0; q < PPL; q++) c1[q] -= temp * c0[q];
        }
        r[0][0] = -ajnorm;
        double ajnorm1 = tb_enorm_warp<PPL, SEQ>(c1, m, lane, 1, wbuf);
        if (ajnorm1 != 0.0) {
            if (tb_elem<PPL>(c1, 1) < 0.0) ajnorm1 = -ajnorm1;
            const double ainv1 = 1.0 / ajnorm1;
#pragma unroll
            for (int q = 0; q < PPL; q++)
                if (lane + 32 * q >= 1) c1[q] = SEQ ? c1[q] / ajnorm1 : c1[q] * ainv1;
            if (lane == 1) c1[0] += 1.0;
        }
        r[1][1] = -ajnorm1;
        r[0][1] = tb_elem<PPL>(c1, 0);
        r[1][0] = 0.0;
        // ---- first iteration: scaling and trust-region radius
        if (iter == 1) {
            for (int j = 0; j < 2; j++) {
                diag[j] = acn[j];
                if (acn[j] == 0.0) diag[j] = 1.0;
                wa3[j] = diag[j] * par[j];
            }
            xnorm = tb_enorm2(wa3[0], wa3[1]);
            delta = factor * xnorm;
            if (delta == 0.0) delta = factor;
        }
        // ---- qtf = first two components of Q^T fvec
#pragma unroll
        for (int q = 0; q < PPL; q++) wa4[q] = fvec[q];
        {
            double a00 = tb_elem<PPL>(c0, 0);
            if (a00 != 0.0) {
                double s = tb_dot_warp<PPL, SEQ>(c0, wa4, m, lane, 0, wbuf);
                double temp = -s / a00;
#pragma unroll
                for (int q = 0; q < PPL; q++) wa4[q] += c0[q] * temp;
            }
            qtf[0] = tb_elem<PPL>(wa4, 0);
            double a11 = tb_elem<PPL>(c1, 1);
            if (a11 != 0.0) {
                double s = tb_dot_warp<PPL, SEQ>(c1, wa4, m, lane, 1, wbuf);
                double temp = -s / a11;
#pragma unroll
                for (int q = 0; q < PPL; q++)
                    if (lane + 32 * q >= 1) wa4[q] += c1[q] * temp;
            }
            qtf[1] = tb_elem<PPL>(wa4, 1);
        }
        // ---- norm of the scaled gradient
        gnorm = 0.0;
        if (fnorm != 0.0) {
            for (int j = 0; j < 2; j++) {
                int l = ipvt[j];
                if (acn[l] != 0.0) {
                    double s = 0.0;
                    for (int i = 0; i <= j; i++) s += r[i][j] * (qtf[i] / fnorm);
                    gnorm = fmax(gnorm, fabs(s / acn[l]));
                }
            }
        }
        if (gnorm <= gtol) {
            info = 4;
            break;
        }
        for (int j = 0; j < 2; j++) diag[j] = fmax(diag[j], acn[j]);
        // ---- inner loop
        bool stop = false;
        for (;;) {
            tb_lmpar2(r, ipvt, diag, qtf, delta, lmp, wa1, wa2);
            for (int j = 0; j < 2; j++) {
                wa1[j] = -wa1[j];
                wa2[j] = par[j] + wa1[j];
                wa3[j] = diag[j] * wa1[j];
            }
            double pnorm = tb_enorm2(wa3[0], wa3[1]);
            if (iter == 1) delta = fmin(delta, pnorm);
            tb_relu_residuals<PPL>(x, y, wa2[0], wa2[1], wa4);
            nfev++;
            double fnorm1 = tb_enorm_warp<PPL, SEQ>(wa4, m, lane, 0, wbuf);
            double actred = -1.0;
            if (p1 * fnorm1 < fnorm) {
                double t = fnorm1 / fnorm;
                actred = 1.0 - t * t;
            }
            for (int j = 0; j < 2; j++) {
                wa3[j] = 0.0;
                double temp = wa1[ipvt[j]];
                for (int i = 0; i <= j; i++) wa3[i] += r[i][j] * temp;
            }
            double temp1 = tb_enorm2(wa3[0], wa3[1]) / fnorm;
            double temp2 = sqrt(lmp) * pnorm / fnorm;
            double prered = temp1 * temp1 + temp2 * temp2 / p5;
            double dirder = -(temp1 * temp1 + temp2 * temp2);
            double ratio = 0.0;
            if (prered != 0.0) ratio = actred / prered;
            if (ratio <= p25) {
                double temp;
                if (actred >= 0.0) temp = p5;
                else temp = p5 * dirder / (dirder + p5 * actred);
                if (p1 * fnorm1 >= fnorm || temp < p1) temp = p1;
                delta = temp * fmin(delta, pnorm / p1);
                lmp /= temp;
            } else if (lmp == 0.0 || ratio >= p75) {
                delta = pnorm / p5;
                lmp = p5 * lmp;
            }
            if (ratio >= p0001) {
                for (int j = 0; j < 2; j++) {
                    par[j] = wa2[j];
                    wa2[j] = diag[j] * par[j];
                }
#pragma unroll
                for (int q = 0; q < PPL; q++) fvec[q] = wa4[q];
                xnorm = tb_enorm2(wa2[0], wa2[1]);
                fnorm = fnorm1;
                iter++;
            }
            if (fabs(actred) <= ftol && prered <= ftol && p5 * ratio <= 1.0) info = 1;
            if (delta <= xtol * xnorm) info = 2;
            if (fabs(actred) <= ftol && prered <= ftol && p5 * ratio <= 1.0 && info == 2) info = 3;
            if (info != 0) { stop = true; break; }
            if (nfev >= maxfev) info = 5;
            if (fabs(actred) <= epsmch && prered <= epsmch && p5 * ratio <= 1.0) info = 6;
            if (delta <= epsmch * xnorm) info = 7;
            if (gnorm <= epsmch) info = 8;
            if (info != 0) { stop = true; break; }
            if (ratio >= p0001) break;
        }
        if (stop) break;
    }
    res.a = par[0];
    res.b = par[1];
    res.info = info;
    res.nfev = nfev;
    // scipy: ier must be in {1,2,3,4}; the covariance needs inv(R) (LAPACK trtri: exact zero on the diagonal -> info > 0)
    // and must be free of NaN, else curve_fit warns (OptimizeWarning) and the reference falls back (:295-299).
    int usable = (info >= 1 && info <= 4) ? 1 : 0;
    if (usable) {
        if (r[0][0] == 0.0 || r[1][1] == 0.0) usable = 0;
        else {
            double i00 = 1.0 / r[0][0], i11 = 1.0 / r[1][1];
            double i01 = -i11 * (i00 * r[0][1]);
            double c00 = i00 * i00 + i01 * i01, c01 = i01 * i11, c11 = i11 * i11;
            if (c00 != c00 || c01 != c01 || c11 != c11) usable = 0;
        }
    }
    res.usable = usable;
}
#endif  // TB_FIRST_GEN_FITS
