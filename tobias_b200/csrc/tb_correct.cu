// tb_correct.cu -- K3b/K3c: per-window relu fits, expected signal, corrected signal, rescaling, verification counts.
//
// Reference: bias_correction, tobias/tools/atacorrect_functions.py:191-410; fast_rolling_math "sum",
// tobias/utils/signals.pyx:149-155 (C float accumulator, :116).  Pipeline of one tb_correct_run:
//   K1   tb_pileup_device      per-strand cut-site counts of every extended region              (:231-243)
//   K3a  tb_score_device       raw bias score of every extended position, both strands          (:254-262)
//   K3b  tb_fit_kernel         one warp per (strand, 100-bp window): LM fit / fallback / empty  (:283-304)
//   K3c  tb_correct_kernel     one block per (region, tile): 10-row mean of the window predictions (:306-309),
//                              rolling sums with the reference's fp32 accumulator semantics (:314-339),
//                              expected / corrected / rescale (:323-350), verification k-mer counts (:358-373),
//                              trimmed per-strand tracks (:381-383)
#include "tb_common.cuh"
#include "tb_lm.cuh"
#include "tb_lm_thread.cuh"

int tb_pileup_device(tb_ctx *ctx, i64 n_regions, const i64 *d_lstart, const i64 *d_lend, const i64 *d_off, i64 total,
                     int shift_fwd, int shift_rev, u32 *d_out);
int tb_score_device(tb_ctx *ctx, const i64 *d_ext_start, const i64 *d_ext_off, i64 n_regions, i64 total, int strand_mask,
                    int nan_invalid, double *d_out_fwd, double *d_out_rev);

#define TB_STEP 10          // window step (atacorrect_functions.py:270)
#define TB_FIT_REC 6        // doubles per window record: mode, a, b, max(prediction), MINPACK info, nfev

struct tb_cor_view {
    const i64 *ext_start;   // linear start of the extended regions
    const i64 *ext_off;     // n+1 offsets into the concatenated extended-position space
    const i64 *out_off;     // n+1 offsets into the concatenated trimmed-output space
    const i64 *win_off;     // n+1 offsets into the window space
    i64 n_regions, ext_total, out_total, n_windows;
    int k_flank, window, L, f_extend;
};

__device__ __forceinline__ int tb_n_windows(i64 reg_len, int k_flank, int window) {
    i64 span = reg_len - 2 * (i64)k_flank - window;       // len(range(k_flank, reg_end - window, 10))
    return span > 0 ? (int)((span + TB_STEP - 1) / TB_STEP) : 0;
}

// ------------------------------------------------------------------------------------------------ K3b
#define TB_FIT_WARPS 8      // warps per block of the fit kernel
#ifndef TB_FIT_MINBLOCKS
#define TB_FIT_MINBLOCKS 2  // resident blocks per SM the register allocation is tuned for
#endif

// SEQ = 1: the m-term sums of the fit run in MINPACK's sequential order (bit-faithful to scipy's lmdif on identical
// inputs, verified window by window in tests/test_lm_exact.py); SEQ = 0: warp-tree sums (faster, ~1 ulp per sum).
template <int PPL, bool SEQ>
__global__ void __launch_bounds__(TB_FIT_WARPS * 32, TB_FIT_MINBLOCKS) tb_fit_kernel(tb_cor_view V, const u32 *__restrict__ pile, const double *__restrict__ biasraw,
                              double *__restrict__ fit) {
    __shared__ double wbuf_all[SEQ ? TB_FIT_WARPS * 32 * PPL : 1];
    double *wbuf = SEQ ? wbuf_all + (threadIdx.x >> 5) * 32 * PPL : wbuf_all;
    const int lane = threadIdx.x & 31;
    const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const i64 n_warps = ((i64)gridDim.x * blockDim.x) >> 5;
    const int m = V.window;
    for (i64 unit = warp; unit < 2 * V.n_windows; unit += n_warps) {
        int strand = unit >= V.n_windows ? 1 : 0;
        i64 wg = unit - (i64)strand * V.n_windows;
        i64 j = tb_upper_bound(V.win_off, V.n_regions + 1, wg) - 1;
        int jw = (int)(wg - V.win_off[j]);
        i64 base = (i64)strand * V.ext_total + V.ext_off[j] + V.k_flank + (i64)TB_STEP * jw;
        double x[PPL], y[PPL];
        double ymax = 0.0;
#pragma unroll
        for (int q = 0; q < PPL; q++) {
            int i = lane + 32 * q;
            x[q] = 0.0;
            y[q] = 0.0;
            if (i < m) {
                x[q] = biasraw[base + i];
                y[q] = (double)pile[base + i];
                ymax = fmax(ymax, y[q]);
            }
        }
        ymax = tb_warp_max(ymax);
        double mode, pa = 0.0, pb = 0.0, pmax = 0.0, pinfo = 0.0, pnfev = 0.0;
        if (ymax > 0.0) {                                          // signalmax > 0 (:290)
            tb_lm_result res;
            tb_lm_fit_warp<PPL, SEQ>(x, y, m, lane, wbuf, res);
            pinfo = (double)res.info;
            pnfev = (double)res.nfev;
            double vmax = -INFINITY;
            if (res.usable) {
                mode = 0.0;
                pa = res.a;
                pb = res.b;
#pragma unroll
                for (int q = 0; q < PPL; q++)
                    if (lane + 32 * q < m) {
                        double t = __dadd_rn(__dmul_rn(pa, x[q]), pb);
                        vmax = fmax(vmax, t > 0.0 ? t : 0.0);
                    }
            } else {                                               // fallback (:296-299)
                mode = 1.0;
                double bmin = INFINITY;
#pragma unroll
                for (int q = 0; q < PPL; q++)
                    if (lane + 32 * q < m && fabs(y[q]) > 1e-8) bmin = fmin(bmin, x[q]);      // ~isclose(signal, 0)
                bmin = tb_warp_min(bmin);
                pb = bmin;
#pragma unroll
                for (int q = 0; q < PPL; q++)
                    if (lane + 32 * q < m) {
                        double t = x[q] - bmin;
                        vmax = fmax(vmax, t < 0.0 ? 0.0 : t);
                    }
            }
            pmax = tb_warp_max(vmax);
        } else {
            mode = 2.0;                                            // np.zeros (:304)
        }
        if (lane == 0) {
            double *o = fit + ((i64)strand * V.n_windows + wg) * TB_FIT_REC;
            o[0] = mode;
            o[1] = pa;
            o[2] = pb;
            o[3] = pmax;
            o[4] = pinfo;
            o[5] = pnfev;
        }
    }
}

// ------------------------------------------------------------------------------------------------ K3b, thread form
// One THREAD per window (tb_lm_thread.cuh).  A block owns a chunk of consecutive windows of one strand (host-built
// table: <= TB_FITT_WIN windows, <= TB_FITT_SEG region segments, <= TB_FIT_Q tile slots), stages their positions once
// into a transposed shared-memory tile (bias f64 + signal u32 = 12 B per position, 120 B per additional window) and
// lets its threads pull windows from a shared counter: a lane that finishes a fit takes the next window at the top of
// the next round, so the 32 lanes of a warp stay busy although fits need 2..18 outer iterations.  Every round runs the
// same phase sequence (QR | trial | finish); lanes skip the phases they do not need.
#ifndef TB_FITT_THREADS
#define TB_FITT_THREADS 256
#endif
#ifndef TB_FITT_WIN
#define TB_FITT_WIN 768
#endif
#define TB_FITT_SEG 32
#ifndef TB_FITT_MINBLOCKS
#define TB_FITT_MINBLOCKS 2
#endif
#define TB_FITT_SMEM ((size_t)10 * TB_FIT_Q * (sizeof(double) + sizeof(u32)) + (size_t)TB_FITT_WIN * sizeof(double))

__global__ void __launch_bounds__(TB_FITT_THREADS, TB_FITT_MINBLOCKS)
    tb_fit_thread_kernel(tb_cor_view V, const i64 *__restrict__ chunk_w0, const u32 *__restrict__ pile,
                         const double *__restrict__ biasraw, double *__restrict__ fit) {
    TB_DYN_SMEM(smem);
    double *tx = (double *)smem;                               // [10][TB_FIT_Q] bias
    double *sf0 = tx + 10 * TB_FIT_Q;                          // [TB_FITT_WIN]  residual norm at p0, -1 = empty window
    u32 *ty = (u32 *)(sf0 + TB_FITT_WIN);                      // [10][TB_FIT_Q] signal
    __shared__ int seg_w0[TB_FITT_SEG + 1], seg_q0[TB_FITT_SEG], seg_np[TB_FITT_SEG];
    __shared__ i64 seg_pos[TB_FITT_SEG];
    __shared__ int n_seg, next;
    const int m = V.window;
    const int strand = blockIdx.y;
    const i64 w_first = chunk_w0[blockIdx.x];
    const int n_win = (int)(chunk_w0[blockIdx.x + 1] - w_first);
    if (threadIdx.x == 0) {
        i64 wg = w_first, wend = w_first + n_win;
        i64 j = tb_upper_bound(V.win_off, V.n_regions + 1, wg) - 1;
        int s = 0, q = 0;
        while (wg < wend && s < TB_FITT_SEG) {
            i64 jw = wg - V.win_off[j];
            i64 hi = V.win_off[j + 1] < wend ? V.win_off[j + 1] : wend;
            int n = (int)(hi - wg);
            if (n > 0) {
                seg_w0[s] = (int)(wg - w_first);
                seg_q0[s] = q;
                seg_pos[s] = (i64)strand * V.ext_total + V.ext_off[j] + V.k_flank + (i64)TB_STEP * jw;
                seg_np[s] = TB_STEP * (n - 1) + m;
                q += (seg_np[s] + 9) / 10;
                s++;
                wg += n;
            }
            j++;
        }
        seg_w0[s] = n_win;
        n_seg = s;
        next = 0;
    }
    __syncthreads();
    for (int s = 0; s < n_seg; s++) {
        const i64 pos = seg_pos[s];
        const int np = seg_np[s], q0 = seg_q0[s];
        for (int lp = threadIdx.x; lp < np; lp += blockDim.x) {
            const int slot = (lp % 10) * TB_FIT_Q + q0 + lp / 10;
            tx[slot] = biasraw[pos + lp];
            ty[slot] = pile[pos + lp];
        }
    }
    __syncthreads();
    // residual norm at p0 and the empty-window test of every window of the chunk, all lanes busy
    for (int cw = threadIdx.x; cw < n_win; cw += blockDim.x) {
        int s = 0;
        while (cw >= seg_w0[s + 1]) s++;
        const int q = seg_q0[s] + (cw - seg_w0[s]);
        sf0[cw] = tb_lmt_fnorm0(tb_win_ref{tx + q, ty + q}, m);
    }
    __syncthreads();

    tb_lmt S;
    tb_win_ref W{tx, ty};
    double *rec = fit;
    bool active = true, need_new = true, need_qr = false;
    while (active) {
        if (need_new) {
            for (;;) {
                const int cw = atomicAdd(&next, 1);
                if (cw >= n_win) {
                    active = false;
                    break;
                }
                rec = fit + ((i64)strand * V.n_windows + w_first + cw) * TB_FIT_REC;
                const double f0 = sf0[cw];
                if (f0 >= 0.0) {                                   // signalmax > 0 (:290)
                    int s = 0;
                    while (cw >= seg_w0[s + 1]) s++;
                    const int q = seg_q0[s] + (cw - seg_w0[s]);
                    W.x = tx + q;
                    W.y = ty + q;
                    tb_lmt_init(S, f0);
                    need_new = false;
                    need_qr = true;
                    break;
                }
                rec[0] = 2.0;                                      // np.zeros (:304)
                rec[1] = rec[2] = rec[3] = rec[4] = rec[5] = 0.0;
            }
        }
        if (active) {
            if (need_qr) {
                tb_lmt_qr(S, W, m);
                need_qr = false;
            }
            if (S.info == 0) need_qr = tb_lmt_trial(S, W, m);
            if (S.info != 0) {
                double mode, pa, pb, pmax;
                tb_lmt_finish(S, W, m, mode, pa, pb, pmax);
                rec[0] = mode;
                rec[1] = pa;
                rec[2] = pb;
                rec[3] = pmax;
                rec[4] = (double)S.info;
                rec[5] = (double)S.nfev;
                need_new = true;
            }
        }
    }
}

// chunk table of tb_fit_thread_kernel: windows [w0[c], w0[c+1]) with bounded window count, segment count and tile slots
static void tb_fit_chunks(const std::vector<i64> &win_off, int window, std::vector<i64> &w0) {
    const int qm = (window + 9) / 10;
    w0.clear();
    w0.push_back(0);
    int nw = 0, nq = 0, ns = 0;
    const i64 n_regions = (i64)win_off.size() - 1;
    for (i64 j = 0; j < n_regions; j++) {
        i64 left = win_off[j + 1] - win_off[j], at = win_off[j];
        while (left > 0) {
            i64 room = TB_FITT_WIN - nw;
            i64 roomq = (i64)TB_FIT_Q - nq - qm + 1;
            if (roomq < room) room = roomq;
            if (room <= 0 || ns == TB_FITT_SEG) {
                w0.push_back(at);
                nw = nq = ns = 0;
                continue;
            }
            i64 n = left < room ? left : room;
            nw += (int)n;
            nq += (int)n - 1 + qm;
            ns++;
            at += n;
            left -= n;
        }
    }
    if (w0.back() != win_off[n_regions]) w0.push_back(win_off[n_regions]);
}

// ------------------------------------------------------------------------------------------------ K3c
// centred rolling sum with the reference's accumulator: valsum is a C float, every += is a double add rounded to float
__device__ __forceinline__ double tb_roll_f32(const double *v, int first, int w) {
    float acc = 0.0f;
    for (int t = 0; t < w; t++) acc = (float)((double)acc + v[first + t]);
    return (double)acc;
}

#define TB_K3C_TILE 512

__global__ void tb_correct_kernel(tb_genome_view g, tb_cor_view V, const int32_t *__restrict__ tile_region,
                                  const int32_t *__restrict__ tile_start, const u32 *__restrict__ pile,
                                  const double *__restrict__ biasraw, const double *__restrict__ fit, double cf,
                                  double *__restrict__ tracks, double *__restrict__ prepost) {
    TB_DYN_SMEM(smem);
    const int w = V.window, k = V.k_flank, L = V.L;
    const int lf = w / 2;
    const int overlaps = w / TB_STEP;
    const i64 j = tile_region[blockIdx.x];
    const int t0 = tile_start[blockIdx.x];
    const int Lr = (int)(V.ext_off[j + 1] - V.ext_off[j]);
    const int nwin = tb_n_windows(Lr, k, w);
    const int t1 = min(Lr, t0 + TB_K3C_TILE);                 // tile = [t0, t1)
    const int c0 = max(0, t0 - w), c1 = min(Lr, t1 + w);      // domain of the raw corrected signal
    const int d0 = max(0, c0 - w), d1 = min(Lr, c1 + w);      // domain of signal / bias prediction
    const int cap = TB_K3C_TILE + 4 * w;
    double *sig = (double *)smem;              // [d0, d1)
    double *bp = sig + cap;                    // [d0, d1)  mean window prediction
    double *unc = bp + cap;                    // [c0, c1)  signal * cf
    double *cor = unc + cap;                   // [c0, c1)  uncorrected - expected
    double *expd = cor + cap;                  // [t0, t1)
    double *ver = expd + cap;                  // [3][4][L] verification accumulators
    const int tid = threadIdx.x, nt = blockDim.x;
    const i64 gext = V.ext_off[j];
    const i64 gstart = V.ext_start[j];

    for (int strand = 0; strand < 2; strand++) {
        const u32 *ps = pile + (i64)strand * V.ext_total + gext;
        const double *br = biasraw + (i64)strand * V.ext_total + gext;
        const double *ft = fit + ((i64)strand * V.n_windows + V.win_off[j]) * TB_FIT_REC;
        for (int i = tid; i < 3 * 4 * L; i += nt) ver[i] = 0.0;
        // ---- signal and bias_prediction = nanmean over the `overlaps` rows (:279-309)
        for (int q = d0 + tid; q < d1; q += nt) {
            sig[q - d0] = (double)ps[q];
            double x = br[q];
            double sum = 0.0;
            int cnt = 0;
            int jmax = q >= k ? (q - k) / TB_STEP : -1;
            if (jmax > nwin - 1) jmax = nwin - 1;
            for (int r = 0; r < overlaps; r++) {
                double val = 0.0;
                bool have = false;
                if (jmax >= r) {
                    int jw = jmax - ((jmax - r) % overlaps);       // last window of row r starting at or before q
                    if (k + TB_STEP * jw + w > q) {
                        have = true;
                        const double *f = ft + (i64)jw * TB_FIT_REC;
                        double mode = f[0], mx = f[3];
                        if (mode == 0.0) {
                            double t = __dadd_rn(__dmul_rn(f[1], x), f[2]);
                            val = t > 0.0 ? t : 0.0;
                            if (mx > 0.0) val = val / mx;
                        } else if (mode == 1.0) {
                            double t = x - f[2];
                            val = t < 0.0 ? 0.0 : t;
                            if (mx > 0.0) val = val / mx;
                        }
                    }
                }
                // rows k_flank.. start as NaN (:280 indexes rows): uncovered cells there do not count
                if (have || r < k) {
                    sum += val;
                    cnt++;
                }
            }
            bp[q - d0] = cnt > 0 ? sum / (double)cnt : (double)NAN;
        }
        __syncthreads();
        // ---- expected and raw corrected signal on [c0, c1) (:314-328)
        for (int q = c0 + tid; q < c1; q += nt) {
            double ssum = 0.0, bsum = 1.0, probas = 0.0;
            if (q >= lf && q - lf + w <= Lr) {
                ssum = tb_roll_f32(sig, q - lf - d0, w);
                double b = tb_roll_f32(bp, q - lf - d0, w);
                bool null = !(fabs(b) > 1e-8);                 // isclose(bias_sum, 0) or NaN
                if (!null) {
                    bsum = b;
                    probas = bp[q - d0] / bsum;
                }
            }
            double e = ssum * probas;
            double u = sig[q - d0] * cf;
            e = e * cf;
            unc[q - c0] = u;
            cor[q - c0] = u - e;
            if (q >= t0 && q < t1) expd[q - t0] = e;
        }
        __syncthreads();
        // ---- rescale (:331-350), outputs, verification counts (:358-373)
        for (int q = t0 + tid; q < t1; q += nt) {
            double c = cor[q - c0];
            double u = unc[q - c0];
            double scale = 1.0;
            if (q >= lf && q - lf + w <= Lr) {
                int first = q - lf - c0;
                float au = 0.0f, aa = 0.0f, ap = 0.0f;
                for (int t = 0; t < w; t++) {
                    double cv = cor[first + t];
                    au = (float)((double)au + unc[first + t]);
                    aa = (float)((double)aa + fabs(cv));
                    ap = (float)((double)ap + (cv < 0.0 ? 0.0 : cv));
                }
                double usum = au, asum = aa, psum = ap;
                double nsum = asum - psum;
                if (psum != 0.0) {
                    scale = (usum - nsum) / psum;
                    if (scale < 1.0) scale = 1.0;
                }
            }
            if (c > 0.0) c = c * scale;
            if (q >= V.f_extend && q < Lr - V.f_extend) {
                i64 o = V.out_off[j] + (q - V.f_extend);
                tracks[((i64)0 * 2 + strand) * V.out_total + o] = u;
                tracks[((i64)1 * 2 + strand) * V.out_total + o] = br[q];
                tracks[((i64)2 * 2 + strand) * V.out_total + o] = expd[q - t0];
                tracks[((i64)3 * 2 + strand) * V.out_total + o] = c;
            }
            if (q > k && q < Lr - k - 1 && (u != 0.0 || c != 0.0)) {
                i64 gp = gstart + q;
                if (tb_nflags(g, gp - k, L) == 0) {
                    u64 kmer = tb_bases(g, gp - k, L);
                    for (int m = 0; m < L; m++) {
                        int s = (int)((kmer >> (2 * m)) & 3u);
                        int mm = m;
                        if (strand) {
                            s ^= 1;
                            mm = L - 1 - m;
                        }
                        if (u > 0.0) atomicAdd(&ver[(0 * 4 + s) * L + mm], u);           // pre_bias counts
                        if (c > 0.0) atomicAdd(&ver[(1 * 4 + s) * L + mm], c);           // post_bias counts
                        else atomicAdd(&ver[(2 * 4 + s) * L + mm], c);                   // post_bias neg_counts
                    }
                }
            }
        }
        __syncthreads();
        for (int i = tid; i < 3 * 4 * L; i += nt) {
            double v = ver[i];
            if (v != 0.0) {
                int which = i / (4 * L), rem = i % (4 * L);
                atomicAdd(&prepost[((i64)strand * 3 + which) * 5 * L + rem], v);
            }
        }
        __syncthreads();
    }
}

// out[i] = tracks[t][0][i] (+ tracks[t][1][i]) as float or double
template <class T>
__global__ void tb_pack_track_kernel(const double *__restrict__ a, const double *__restrict__ b, i64 n, T *__restrict__ out) {
    i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (T)(b ? a[i] + b[i] : a[i]);
}

// ---------------------------------------------------------------------------------------------- host side
extern "C" int tb_correct_run(tb_ctx *ctx, int64_t n_regions, const int32_t *contig, const int64_t *start, const int64_t *end,
                              int k_flank, int window, double correction_factor, int shift_fwd, int shift_rev,
                              int lm_exact_order) {
    if (!ctx) return TB_ERR_ARG;
    ctx->cor.valid = false;
    if (window < TB_STEP || window > 512) return tb_fail(ctx, TB_ERR_ARG, "tb_correct_run: window must be in [10, 512]");
    if (ctx->model[0].is_dwm < 0 || ctx->model[1].is_dwm < 0)
        return tb_fail(ctx, TB_ERR_STATE, "tb_correct_run: call tb_set_bias_model for both strands first");
    const int L = ctx->model[0].L;
    if (L != 2 * k_flank + 1) return tb_fail(ctx, TB_ERR_ARG, "tb_correct_run: k_flank %d does not match the model length %d", k_flank, L);
    std::vector<i64> ls, le;
    TB_TRY(tb_linear_regions(ctx, n_regions, contig, start, end, ls, le, false));
    const int f_extend = k_flank + window / 2;
    auto &C = ctx->cor;
    C.n_regions = n_regions;
    C.k_flank = k_flank;
    C.window = window;
    C.L = L;
    C.ext_start.resize(n_regions);
    C.ext_off.assign(n_regions + 1, 0);
    C.out_off.assign(n_regions + 1, 0);
    C.win_off.assign(n_regions + 1, 0);
    std::vector<i64> ext_end(n_regions);
    std::vector<int32_t> tile_region, tile_start;
    for (i64 i = 0; i < n_regions; i++) {
        int c = contig[i];
        if (start[i] - f_extend < 0 || end[i] + f_extend > ctx->contig_len[c])
            return tb_fail(ctx, TB_ERR_ARG, "tb_correct_run: region %lld extended by %d leaves contig %d", i, f_extend, c);
        C.ext_start[i] = ls[i] - f_extend;
        ext_end[i] = le[i] + f_extend;
        i64 reg_len = ext_end[i] - C.ext_start[i];
        if (reg_len > ((i64)1 << 30)) return tb_fail(ctx, TB_ERR_ARG, "tb_correct_run: region %lld too long", i);
        i64 span = reg_len - 2 * (i64)k_flank - window;
        i64 nw = span > 0 ? (span + TB_STEP - 1) / TB_STEP : 0;
        if (nw == 0) return tb_fail(ctx, TB_ERR_ARG, "tb_correct_run: region %lld holds no window (reference behaviour undefined)", i);
        C.ext_off[i + 1] = C.ext_off[i] + reg_len;
        C.out_off[i + 1] = C.out_off[i] + (le[i] - ls[i]);
        C.win_off[i + 1] = C.win_off[i] + nw;
        for (i64 t = 0; t < reg_len; t += TB_K3C_TILE) {
            tile_region.push_back((int32_t)i);
            tile_start.push_back((int32_t)t);
        }
    }
    C.ext_total = C.ext_off[n_regions];
    C.out_total = C.out_off[n_regions];
    C.n_windows = C.win_off[n_regions];
    if (n_regions == 0) {
        C.valid = true;
        return TB_OK;
    }
    TB_TRY(tb_h2d(ctx, C.d_ext_start, C.ext_start.data(), sizeof(i64) * n_regions));
    TB_TRY(tb_h2d(ctx, C.d_ext_end, ext_end.data(), sizeof(i64) * n_regions));
    TB_TRY(tb_h2d(ctx, C.d_ext_off, C.ext_off.data(), sizeof(i64) * (n_regions + 1)));
    TB_TRY(tb_h2d(ctx, C.d_out_off, C.out_off.data(), sizeof(i64) * (n_regions + 1)));
    TB_TRY(tb_h2d(ctx, C.d_win_off, C.win_off.data(), sizeof(i64) * (n_regions + 1)));
    TB_TRY(tb_h2d(ctx, ctx->reg_a, tile_region.data(), sizeof(int32_t) * tile_region.size()));
    TB_TRY(tb_h2d(ctx, ctx->reg_b, tile_start.data(), sizeof(int32_t) * tile_start.size()));
    TB_TRY(tb_reserve(ctx, C.pile, (size_t)C.ext_total * 2 * sizeof(u32)));
    TB_TRY(tb_reserve(ctx, C.biasraw, (size_t)C.ext_total * 2 * sizeof(double)));
    TB_TRY(tb_reserve(ctx, C.fit, (size_t)C.n_windows * 2 * TB_FIT_REC * sizeof(double)));
    TB_TRY(tb_reserve(ctx, C.tracks, (size_t)C.out_total * 8 * sizeof(double)));
    TB_TRY(tb_reserve(ctx, C.prepost, (size_t)2 * 3 * 5 * L * sizeof(double)));
    TB_CUDA(ctx, cudaMemsetAsync(C.prepost.p, 0, (size_t)2 * 3 * 5 * L * sizeof(double), ctx->stream));

    tb_timer_start(ctx);
    tb_mark(ctx, 0);
    TB_TRY(tb_pileup_device(ctx, n_regions, C.d_ext_start.as<i64>(), C.d_ext_end.as<i64>(), C.d_ext_off.as<i64>(), C.ext_total,
                            shift_fwd, shift_rev, C.pile.as<u32>()));
    tb_mark(ctx, 1);
    TB_TRY(tb_score_device(ctx, C.d_ext_start.as<i64>(), C.d_ext_off.as<i64>(), n_regions, C.ext_total, 3, 0,
                           C.biasraw.as<double>(), C.biasraw.as<double>() + C.ext_total));
    tb_mark(ctx, 2);
    tb_cor_view V{C.d_ext_start.as<i64>(), C.d_ext_off.as<i64>(), C.d_out_off.as<i64>(), C.d_win_off.as<i64>(),
                  n_regions, C.ext_total, C.out_total, C.n_windows, k_flank, window, L, f_extend};
    if (lm_exact_order < 2 && window <= 10 * (TB_FIT_Q - 2)) {
        // thread-per-window fits in MINPACK's own summation order (bit-faithful to scipy on identical input)
        std::vector<i64> w0;
        tb_fit_chunks(C.win_off, window, w0);
        TB_TRY(tb_h2d(ctx, ctx->reg_c, w0.data(), sizeof(i64) * w0.size()));
        const size_t smem = TB_FITT_SMEM;
        auto kf = tb_fit_thread_kernel;
        TB_CUDA(ctx, cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TB_LAUNCH(ctx, kf, dim3((unsigned)(w0.size() - 1), 2), TB_FITT_THREADS, smem, V, (const i64 *)ctx->reg_c.as<i64>(),
                  (const u32 *)C.pile.as<u32>(), (const double *)C.biasraw.as<double>(), C.fit.as<double>());
        TB_TRY(tb_check(ctx, cudaGetLastError(), "tb_fit_thread_kernel"));
    } else {
        const int seq_sums = lm_exact_order != 2;       // warp-per-window kernels: 2 = tree sums, 3 = MINPACK-order sums
        int block = TB_FIT_WARPS * 32;
        i64 want = tb_div_up(2 * C.n_windows, block / 32), capg = (i64)ctx->sm_count * 8;
        unsigned grid = (unsigned)(want < capg ? want : capg);
#define TB_FIT_CASE(PPL, SEQ)                                                                                             \
    {                                                                                                                     \
        auto kf = tb_fit_kernel<PPL, SEQ>;                                                                                \
        TB_LAUNCH(ctx, kf, grid, block, 0, V, (const u32 *)C.pile.as<u32>(), (const double *)C.biasraw.as<double>(),      \
                  C.fit.as<double>());                                                                                    \
    }
        if (seq_sums) {
            if (window <= 128) TB_FIT_CASE(4, true) else if (window <= 256) TB_FIT_CASE(8, true) else TB_FIT_CASE(16, true)
        } else {
            if (window <= 128) TB_FIT_CASE(4, false) else if (window <= 256) TB_FIT_CASE(8, false) else TB_FIT_CASE(16, false)
        }
#undef TB_FIT_CASE
        TB_TRY(tb_check(ctx, cudaGetLastError(), "tb_fit_kernel"));
    }
    tb_mark(ctx, 3);
    {
        tb_genome_view g{ctx->seq2.as<u32>(), ctx->nmask.as<u32>(), ctx->genome_bases};
        size_t smem = ((size_t)5 * (TB_K3C_TILE + 4 * window) + 3 * 4 * L) * sizeof(double);
        auto kc = tb_correct_kernel;
        TB_CUDA(ctx, cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TB_LAUNCH(ctx, kc, (unsigned)tile_region.size(), 256, smem, g, V, (const int32_t *)ctx->reg_a.as<int32_t>(),
                  (const int32_t *)ctx->reg_b.as<int32_t>(), (const u32 *)C.pile.as<u32>(), (const double *)C.biasraw.as<double>(),
                  (const double *)C.fit.as<double>(), correction_factor, C.tracks.as<double>(), C.prepost.as<double>());
        TB_TRY(tb_check(ctx, cudaGetLastError(), "tb_correct_kernel"));
    }
    tb_mark(ctx, 4);
    tb_timer_stop(ctx);
    tb_marks_collect(ctx, 5);      // stage_ms: [0] K1 pileup, [1] K3a score, [2] K3b fit, [3] K3c correct
    TB_TRY(tb_sync(ctx, "tb_correct_run"));
    C.valid = true;
    return TB_OK;
}

extern "C" int tb_correct_download(tb_ctx *ctx, int split_strands, int as_f64, void *tracks[4][2], double *prepost_out) {
    if (!ctx || !tracks) return TB_ERR_ARG;
    auto &C = ctx->cor;
    if (!C.valid) return tb_fail(ctx, TB_ERR_STATE, "tb_correct_download: no completed tb_correct_run");
    const size_t esz = as_f64 ? 8 : 4;
    const i64 n = C.out_total;
    if (n > 0) {
        TB_TRY(tb_reserve(ctx, ctx->scratch0, (size_t)n * esz));
        for (int t = 0; t < 4; t++)
            for (int s = 0; s < 2; s++) {
                void *dst = tracks[t][s];
                if (!dst) continue;
                if (!split_strands && s == 1) return tb_fail(ctx, TB_ERR_ARG, "tb_correct_download: tracks[t][1] must be NULL unless split_strands");
                const double *a = C.tracks.as<double>() + ((i64)t * 2 + (split_strands ? s : 0)) * n;
                const double *b = split_strands ? nullptr : C.tracks.as<double>() + ((i64)t * 2 + 1) * n;
                i64 want = tb_div_up(n, 256), capg = (i64)ctx->sm_count * 16;
                unsigned grid = (unsigned)(want < capg ? want : capg);
                if (as_f64) {
                    auto kp = tb_pack_track_kernel<double>;
                    TB_LAUNCH(ctx, kp, grid, 256, 0, a, b, n, ctx->scratch0.as<double>());
                } else {
                    auto kp = tb_pack_track_kernel<float>;
                    TB_LAUNCH(ctx, kp, grid, 256, 0, a, b, n, ctx->scratch0.as<float>());
                }
                TB_TRY(tb_d2h(ctx, dst, ctx->scratch0.p, (size_t)n * esz));
            }
    }
    if (prepost_out) TB_TRY(tb_d2h(ctx, prepost_out, C.prepost.p, (size_t)2 * 3 * 5 * C.L * sizeof(double)));
    return tb_sync(ctx, "tb_correct_download");
}

extern "C" int tb_correct_fit_trace(tb_ctx *ctx, int64_t *n_windows_total, double *out) {
    if (!ctx) return TB_ERR_ARG;
    auto &C = ctx->cor;
    if (!C.valid) return tb_fail(ctx, TB_ERR_STATE, "tb_correct_fit_trace: no completed tb_correct_run");
    if (n_windows_total) *n_windows_total = C.n_windows;
    if (out && C.n_windows) TB_TRY(tb_d2h(ctx, out, C.fit.p, (size_t)C.n_windows * 2 * TB_FIT_REC * sizeof(double)));
    return TB_OK;
}
