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
#include <chrono>
#include "tb_lm.cuh"
#include "tb_lm_thread.cuh"

int tb_pileup_device(tb_ctx *ctx, i64 n_regions, const i64 *d_lstart, const i64 *d_lend, const i64 *d_off, i64 total,
                     int shift_fwd, int shift_rev, u32 *d_out, int list_semantics = 0);
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

#ifdef TB_FIRST_GEN_FITS      // first-generation warp-per-window fit kernel (lm_exact_order 2 / 3): test / comparison builds only
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

#endif  // TB_FIRST_GEN_FITS

// ------------------------------------------------------------------------------------------------ K3b, round form
// One THREAD per window (tb_lm_thread.cuh); the fits advance in ROUNDS over the whole launch: the state of a
// fit (16 doubles + 4 ints) lives in HBM between rounds and every round runs over compacted lists of the windows
// that still need work -- list Q (next step: Jacobian/QR + trial) and list T (trial only, after a rejected step) -- so
// the 32 lanes of every warp always execute the same phase on live windows; nothing waits for a neighbour's longer
// fit.  The window data are read through L1 from a transposed copy (plane v = position mod 10, slot q = window index
// + 9 per preceding region), in which the lanes of a warp (consecutive windows) read consecutive addresses.
#define TB_FITG_NST 17

struct tb_fitg_view {
    double *st;             // [TB_FITG_NST][n_slots] fit state, structure of arrays
    int4 *sti;              // [n_slots] ipvt0, iter, nfev, info
    u32 *wq;                // [n_slots] tile slot q of each window
    const double *gx;       // [10][qtot] bias
    const u32 *gy;          // [10][qtot] signal
    i64 qtot, n_slots;
};

// TYY: the moment form carries one more constant of the fit (17th plane); the other forms neither load nor store it
template <bool TYY = false>
__device__ __forceinline__ void tb_fitg_load(const tb_fitg_view &G, i64 slot, tb_lmt &S) {
    const double *p = G.st + slot;
    const i64 n = G.n_slots;
    S.par0 = p[0 * n]; S.par1 = p[1 * n]; S.fnorm = p[2 * n]; S.diag0 = p[3 * n]; S.diag1 = p[4 * n]; S.delta = p[5 * n];
    S.lmp = p[6 * n]; S.xnorm = p[7 * n]; S.gnorm = p[8 * n]; S.r00 = p[9 * n]; S.r01 = p[10 * n]; S.r11 = p[11 * n];
    S.qtf0 = p[12 * n]; S.qtf1 = p[13 * n]; S.acn0 = p[14 * n]; S.acn1 = p[15 * n];
    S.tyy = TYY ? p[16 * n] : 0.0;
    int4 v = G.sti[slot];
    S.ipvt0 = v.x; S.iter = v.y; S.nfev = v.z; S.info = v.w;
}
template <bool TYY = false>
__device__ __forceinline__ void tb_fitg_store(const tb_fitg_view &G, i64 slot, const tb_lmt &S) {
    double *p = G.st + slot;
    const i64 n = G.n_slots;
    p[0 * n] = S.par0; p[1 * n] = S.par1; p[2 * n] = S.fnorm; p[3 * n] = S.diag0; p[4 * n] = S.diag1; p[5 * n] = S.delta;
    p[6 * n] = S.lmp; p[7 * n] = S.xnorm; p[8 * n] = S.gnorm; p[9 * n] = S.r00; p[10 * n] = S.r01; p[11 * n] = S.r11;
    p[12 * n] = S.qtf0; p[13 * n] = S.qtf1; p[14 * n] = S.acn0; p[15 * n] = S.acn1;
    if (TYY) p[16 * n] = S.tyy;
    G.sti[slot] = int4{S.ipvt0, S.iter, S.nfev, S.info};
}
// warp-aggregated append of `slot` to list[ctr] for the lanes with `want`; every lane of the warp must call it
__device__ __forceinline__ void tb_fitg_append(bool want, u32 slot, u32 *__restrict__ list, unsigned long long *ctr) {
    const unsigned FULL = 0xffffffffu;
    const unsigned mask = __ballot_sync(FULL, want);
    if (!mask) return;
    const int lane = threadIdx.x & 31, leader = __ffs((int)mask) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(ctr, (unsigned long long)__popc(mask));
    base = __shfl_sync(FULL, base, leader);
    if (want) list[base + __popc(mask & ((1u << lane) - 1u))] = slot;
}

// transposed copy of (bias, signal) of every window position; one block per (region, strand)
__global__ void tb_fitg_transpose_kernel(tb_cor_view V, const u32 *__restrict__ pile, const double *__restrict__ biasraw, i64 qs,
                                         i64 qtot, int qm, double *__restrict__ gx, u32 *__restrict__ gy, int32_t *__restrict__ slot_win) {
    const i64 j = blockIdx.x;
    const int strand = blockIdx.y;
    const i64 nw = V.win_off[j + 1] - V.win_off[j];
    if (nw <= 0) return;
    const i64 q0 = (i64)strand * qs + V.win_off[j] + (i64)(qm - 1) * j;
    if (slot_win)          // window (strand * n_windows + index) of every slot of this region, -1 for the qm - 1 slots that only hold window tails
        for (int u = threadIdx.x; u < nw + qm - 1; u += blockDim.x)
            slot_win[q0 + u] = u < nw ? (int32_t)((i64)strand * V.n_windows + V.win_off[j] + u) : -1;
    const i64 pos = (i64)strand * V.ext_total + V.ext_off[j] + V.k_flank;
    const int np = (int)(TB_STEP * (nw - 1)) + V.window;
    for (int lp = threadIdx.x; lp < np; lp += blockDim.x) {
        const i64 at = (i64)(lp % 10) * qtot + q0 + lp / 10;
        gx[at] = biasraw[pos + lp];
        gy[at] = pile[pos + lp];
    }
}

// per window: tile slot, residual norm at p0, empty-window record, initial state, first Q list
// FAST_HEAD (one-pass forms; 1: forward-difference inner products, 2: moments of the active set): the same pass also runs the first
// outer-iteration head, the window goes to the trial-only list
template <int FAST_HEAD>
__global__ void tb_fitg_init_kernel(tb_cor_view V, tb_fitg_view G, i64 qs, int qm, double *__restrict__ fit, u32 *__restrict__ listq,
                                    unsigned long long *ctr) {
    const i64 n_round = ((G.n_slots + 31) / 32) * 32;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 slot = (i64)blockIdx.x * blockDim.x + threadIdx.x; slot < n_round; slot += stride) {
        bool live = false;
        if (slot < G.n_slots) {
            const int strand = slot >= V.n_windows ? 1 : 0;
            const i64 wg = slot - (i64)strand * V.n_windows;
            const i64 j = tb_upper_bound(V.win_off, V.n_regions + 1, wg) - 1;
            const i64 q = (i64)strand * qs + wg + (i64)(qm - 1) * j;
            G.wq[slot] = (u32)q;
            tb_win_stats st;
            tb_lmt S;
            const tb_win_ref W{G.gx + q, G.gy + q, G.qtot};
            const double f0 = FAST_HEAD == 2 ? tb_lmt_start_moments(S, W, V.window, st)
                                             : (FAST_HEAD == 1 ? tb_lmt_start_fast(S, W, V.window, st) : tb_lmt_fnorm0_stats(W, V.window, st));
            if (f0 >= 0.0) {                                           // signalmax > 0 (:290)
                if (!FAST_HEAD) tb_lmt_init(S, f0);
                tb_fitg_store<FAST_HEAD == 2>(G, slot, S);
                double *rec = fit + slot * TB_FIT_REC;                 // parked in the window's record until its fit ends
                rec[1] = st.xmax;
                rec[2] = st.xmin;
                rec[3] = st.bmin;
                live = true;
            } else {
                double *rec = fit + slot * TB_FIT_REC;
                rec[0] = 2.0;                                          // np.zeros (:304)
                rec[1] = rec[2] = rec[3] = rec[4] = rec[5] = 0.0;
            }
        }
        tb_fitg_append(live, (u32)slot, listq, ctr);
    }
}

// one round over a list: WITH_QR: outer-iteration head + trial; else trial only.  ctr[0] / ctr[1]: next Q / T list sizes
#ifndef TB_FITG_MINBLOCKS
#define TB_FITG_MINBLOCKS 4
#endif
#ifndef TB_FITG_PREFETCH_FAST
#define TB_FITG_PREFETCH_FAST 0      // 1: one-pass form prefetches both ends of the warp's span in every plane at the start of a round (measured: 58.3 -> 64.5 ms, off)
#endif
#ifndef TB_FITG_FUSE
#define TB_FITG_FUSE 1      // 0: separate Jacobian and trial passes in the one-pass form (comparison builds)
#endif
// One-pass form: five blocks per SM (96 registers) measured 58.2 ms on hg38 against 59.8 with four (128) and 59.5 with six (80).
// FAST: 0 MINPACK-exact arithmetic, 1 one-pass form with forward-difference inner products, 2 one-pass form with the moments of the active set
template <bool WITH_QR, int FAST>
__global__ void __launch_bounds__(128, FAST ? TB_FITG_MINBLOCKS + 1 : TB_FITG_MINBLOCKS)
    tb_fitg_round_kernel(tb_fitg_view G, int m, const u32 *__restrict__ list, i64 n_list, double *__restrict__ fit,
                         u32 *__restrict__ nextq, u32 *__restrict__ nextt, unsigned long long *ctr) {
    const i64 n_round = ((n_list + 31) / 32) * 32;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < n_round; k += stride) {
        bool to_q = false, to_t = false;
        u32 slot = 0;
        if (k < n_list) {
            slot = list[k];
            tb_lmt S;
            tb_fitg_load<FAST == 2>(G, slot, S);
            const i64 q = G.wq[slot];
            const tb_win_ref W{G.gx + q, G.gy + q, G.qtot};
#ifndef TB_EMU
            if (!FAST || TB_FITG_PREFETCH_FAST) {   // (measured: -5 % on the exact form)
                // The lanes of a warp read consecutive slots q .. q + 31 of a plane and move up by one slot every ten points, so the top
                // of the warp's span enters the cache one sector at a time, each time as a miss the pass loop has to wait for.  Ask for
                // the last slot every lane will need in every plane now: the whole span of the round is then in flight at once.
                const int top = (m - 1) / 10;
#pragma unroll
                for (int v = 0; v < 10; v++) {
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(W.x + (i64)v * W.stride + top));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(W.y + (i64)v * W.stride + top));
                    if (FAST) {        // plane-major traversal: the first touch of every plane would otherwise be a miss two points ahead of its use
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(W.x + (i64)v * W.stride));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(W.y + (i64)v * W.stride));
                    }
                }
            }
#endif
            if (WITH_QR) {
                if (FAST == 2) tb_lmt_qr_moments(S, W, m);
                else if (FAST == 1) tb_lmt_qr_fast(S, W, m);
                else tb_lmt_qr(S, W, m);
            }
            // one-pass forms: fused trial + next head, so every surviving window continues on the trial-only list
            bool accepted = false;
            if (S.info == 0) accepted = tb_lmt_trial<FAST != 0, FAST && TB_FITG_FUSE, FAST == 2>(S, W, m);
            if (S.info != 0) {
                double mode, pa, pb, pmax;
                double *rec = fit + (i64)slot * TB_FIT_REC;
                const tb_win_stats st{rec[1], rec[2], rec[3]};         // left there by tb_fitg_init_kernel
                tb_lmt_finish_stats(S, st, mode, pa, pb, pmax);
                rec[0] = mode;
                rec[1] = pa;
                rec[2] = pb;
                rec[3] = pmax;
                rec[4] = (double)S.info;
                rec[5] = (double)S.nfev;
            } else {
                tb_fitg_store<FAST == 2>(G, slot, S);
                to_q = accepted && !(FAST == 2 || (FAST && TB_FITG_FUSE));
                to_t = !to_q;
            }
        }
        tb_fitg_append(to_q, slot, nextq, ctr);
        tb_fitg_append(to_t, slot, nextt, ctr + 1);
    }
}

// Tail of the rounds: once few windows are left a round costs the latency of one thread's iteration plus a launch and a host
// read-back, whatever the count.  The remaining windows are then finished in ONE launch, every thread iterating its window to the end
// (lanes diverge, but the GPU is almost idle at that point anyway).
template <int FAST>
__global__ void __launch_bounds__(128, TB_FITG_MINBLOCKS)
    tb_fitg_finish_kernel(tb_fitg_view G, int m, const u32 *__restrict__ listq, i64 nq, const u32 *__restrict__ listt, i64 nt,
                          double *__restrict__ fit) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < nq + nt; k += stride) {
        const u32 slot = k < nq ? listq[k] : listt[k - nq];
        bool need_qr = k < nq;
        tb_lmt S;
        tb_fitg_load<FAST == 2>(G, slot, S);
        const i64 q = G.wq[slot];
        const tb_win_ref W{G.gx + q, G.gy + q, G.qtot};
        for (;;) {
            if (need_qr) {
                if (FAST == 2) tb_lmt_qr_moments(S, W, m);
                else if (FAST == 1) tb_lmt_qr_fast(S, W, m);
                else tb_lmt_qr(S, W, m);
            }
            if (S.info == 0) {
                const bool accepted = tb_lmt_trial<FAST != 0, false, FAST == 2>(S, W, m);
                need_qr = accepted && FAST != 2;          // moment form: the next head ran inside the trial
            }
            if (S.info != 0) break;
        }
        double mode, pa, pb, pmax;
        double *rec = fit + (i64)slot * TB_FIT_REC;
        const tb_win_stats st{rec[1], rec[2], rec[3]};
        tb_lmt_finish_stats(S, st, mode, pa, pb, pmax);
        rec[0] = mode;
        rec[1] = pa;
        rec[2] = pb;
        rec[3] = pmax;
        rec[4] = (double)S.info;
        rec[5] = (double)S.nfev;
    }
}

// ------------------------------------------------------------------------------------------------ K3b, persistent form (one-pass fits)
// The round form above sends the state of every fit (144 B) through HBM once per outer iteration and reads the window data through L1
// from wherever the compacted lists point (one sector per lane once the lists thin out).  Here a block stages a TILE of consecutive
// slots of the transposed window data in shared memory (10 planes x (TS + 12) slots x 12 B) and its threads fit the windows of the tile
// to the end with the state in registers: lane l of every warp takes the slots = l (mod 32) of the tile from a shared counter of that
// residue class, so at any moment the 32 lanes of a warp read 32 different banks whatever windows they hold (the slots a lane reads
// stay = l + u mod 32 for every lane at point u of a plane), and the 16 lanes sharing a class balance each other.  One loop, one call
// site of the step function: lanes at different stages of different fits share every pass.  A fit that needs more than `cap` passes
// is parked (state to HBM, slot appended to the trial list) and finished by the round kernels, so no tile waits for a long fit.
struct tb_fitp_tile {
    int ts, tsp;               // slots per tile, plane stride in shared memory (ts + 12)
};

template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
    tb_fitp_kernel(tb_fitg_view G, const int32_t *__restrict__ slot_win, tb_fitp_tile T, i64 n_tiles, int m, int cap,
                   double *__restrict__ fit, u32 *__restrict__ listt, unsigned long long *ctr) {
    TB_DYN_SMEM(smem);
    double *sx = (double *)smem;                                // [10][tsp]
    u32 *sy = (u32 *)(sx + (size_t)10 * T.tsp);                 // [10][tsp]
    __shared__ u32 qctr[32];
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned FULL = 0xffffffffu;
    for (i64 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const i64 t0 = tile * T.ts;
        __syncthreads();                                        // every fit of the previous tile has ended
        {   // stage the tile: 16-byte loads (t0, tsp and the plane stride in HBM are multiples of 4)
            const int nx = T.tsp / 2, ny = T.tsp / 4;
            for (int i = tid; i < 10 * nx; i += NT) {
                const int v = i / nx, u = i - v * nx;
                ((double2 *)(sx + (size_t)v * T.tsp))[u] = ((const double2 *)(G.gx + (i64)v * G.qtot + t0))[u];
            }
            for (int i = tid; i < 10 * ny; i += NT) {
                const int v = i / ny, u = i - v * ny;
                ((uint4 *)(sy + (size_t)v * T.tsp))[u] = ((const uint4 *)(G.gy + (i64)v * G.qtot + t0))[u];
            }
            if (tid < 32) qctr[tid] = 0u;
        }
        __syncthreads();
        tb_lmt S;
        tb_win_stats st{0.0, 0.0, 0.0};
        int sl = 0, passes = 0, win = -1;
        bool busy = false, first = false, drained = false;
        for (;;) {
            if (!busy && !drained) {                            // take the next slot of this lane's residue class
                for (;;) {
                    const u32 k = atomicAdd(&qctr[lane], 1u);
                    sl = lane + 32 * (int)k;
                    if (sl >= T.ts) {
                        drained = true;
                        break;
                    }
                    win = slot_win[t0 + sl];
                    if (win >= 0) {
                        busy = true;
                        first = true;
                        passes = 0;
                        break;
                    }
                }
            }
            if (!__any_sync(FULL, busy)) break;
            bool parked = false;
            if (busy) {
                const tb_win_ref W{sx + sl, sy + sl, (i64)T.tsp};
                const bool have = tb_lmt_step_fast(S, st, W, m, first);
                first = false;
                passes++;
                double *rec = fit + (i64)win * TB_FIT_REC;
                if (!have) {                                    // np.zeros (:304)
                    rec[0] = 2.0;
                    rec[1] = rec[2] = rec[3] = rec[4] = rec[5] = 0.0;
                    busy = false;
                } else if (S.info != 0) {
                    double mode, pa, pb, pmax;
                    tb_lmt_finish_stats(S, st, mode, pa, pb, pmax);
                    rec[0] = mode;
                    rec[1] = pa;
                    rec[2] = pb;
                    rec[3] = pmax;
                    rec[4] = (double)S.info;
                    rec[5] = (double)S.nfev;
                    busy = false;
                } else if (passes >= cap) {                     // park: the round kernels finish it
                    tb_fitg_store(G, win, S);
                    G.wq[win] = (u32)(t0 + sl);
                    rec[1] = st.xmax;
                    rec[2] = st.xmin;
                    rec[3] = st.bmin;
                    parked = true;
                    busy = false;
                }
            }
            tb_fitg_append(parked, (u32)win, listt, ctr + 1);
        }
    }
}

// ------------------------------------------------------------------------------------------------ K3c
// centred rolling sum with the reference's accumulator: valsum is a C float, every += is a double add rounded to float.
// The accumulator is kept as a double that always holds a float value; the rounding to float after each add is done in the
// fp64 pipe instead of two trips through the conversion unit: with E the exponent of s (at least -126: below that the float
// grid is the fixed denormal grid 2^-149), adding and subtracting c = 1.5 * 2^(E + 29) rounds s to a multiple of 2^(E - 23)
// = the float grid at s, to nearest / ties to even by the hardware's own rounding of s + c (c is a multiple of 2^(E + 28),
// so evenness on the grid is preserved).  NaN and infinity pass through; sums here never approach the float overflow.
__device__ __forceinline__ double tb_round_f32(double s) {
#ifdef TB_K3C_CVT                                  // comparison build: through the conversion unit
    return (double)(float)s;
#endif
    int e = __double2hiint(s) & 0x7ff00000;
    e = max(e, (1023 - 126) << 20);
    const double c = __hiloint2double(e + ((29 << 20) | 0x00080000), 0);
    return __dadd_rn(__dadd_rn(s, c), -c);
}

__device__ __forceinline__ double tb_roll_f32(const double *v, int first, int w) {
    double acc = 0.0;
    for (int t = 0; t < w; t++) {
        const double x = v[first + t];
        if (x != 0.0) acc = tb_round_f32(__dadd_rn(acc, x));   // adding 0.0 never changes the accumulator
    }
    return acc;
}

#define TB_K3C_TILE_MAX 1024     // positions per tile (chosen per launch from the region lengths, a multiple of 64 up to this)
#define TB_K3C_THREADS 256
#define TB_K3C_FREC 6            // doubles per staged fit record: mode, max(prediction), a, b, 1 / max, pad (16-byte aligned pairs)

// Persistent blocks walk the (region, 512-bp tile) list.  Per tile and strand:
//   1. signal + bias prediction (mean of the 10 window predictions) over the tile widened by 2 * window; integer prefix
//      sums of the signal and of its non-zero flags, list of the non-zero positions
//   2. expected / raw corrected over the tile widened by window.  rolling sum of the signal: the reference's float
//      accumulator is exact on integers below 2^24, so the integer prefix difference is the same number (sequential loop
//      beyond that); rolling sum of the bias prediction: sequential float-rounded accumulation, as the reference
//   3. rescale factor and outputs.  sum of uncorrected = signal * cf: adding 0.0 never changes the accumulator, so only
//      the non-zero signal positions are accumulated, in order; |corrected| and corrected+ sums: sequential
//   4. verification counts: pre / post(+) / post(-) weights of every valid position are left in shared memory; lane m of
//      a warp then owns k-mer offset m and adds the non-zero weights of the warp's strip of positions into 12 private (weight kind,
//      base) accumulators -- no atomics in the loop (first version: 25 x 3 shared-memory fp64 atomics = CAS loops per position)
// FAST (lm_exact_order = 0, the DWM default: a tolerance path anyway): the three float-accumulator rolling sums of a position (bias
// prediction, |corrected|, corrected+) are differences of block-wide fp64 prefix sums, rounded to float once (the reference's 100
// sequential float roundings scatter ~3e-7 around that value); the uncorrected sum is cf * (integer prefix difference).  O(1) per position
// and sum instead of 100 dependent add + round steps.  Shared-memory layout per mode: tb_k3c_layout.
struct tb_k3c_layout {
    int capd;        // doubles per position array (cap + 1: room for the total of a prefix sum)
    int off_ver, off_psum, off_pcnt, off_wsum, off_nzpos, off_wpre, off_bases, off_frec, total;     // byte offsets
};
static inline tb_k3c_layout tb_k3c_make_layout(int tile_len, int w, int L, bool fast) {
    tb_k3c_layout y;
    const int cap = tile_len + 4 * w;
    y.capd = cap + 1;
    size_t o = (size_t)5 * y.capd * sizeof(double);
    y.off_ver = (int)o;
    o += (size_t)2 * 3 * 4 * L * sizeof(double);
    y.off_psum = (int)o;
    o += (size_t)(cap + 1) * sizeof(u32);
    y.off_wsum = (int)o;
    o += 64 * sizeof(u32);
    o = (o + 7) & ~(size_t)7;
    y.off_pcnt = y.off_nzpos = y.off_wpre = (int)o;
    if (fast) {
        o += (size_t)tile_len * sizeof(double);                     // pre weights of phase 4 (the exact form keeps them in `sig`)
    } else {
        o += (size_t)(cap + 1) * sizeof(u32);
        y.off_nzpos = (int)o;
        o += (size_t)cap * sizeof(unsigned short);
    }
    y.off_bases = (int)o;
    o += (size_t)tile_len + L + 16;
    o = (o + 15) & ~(size_t)15;
    y.off_frec = (int)o;
    o += (size_t)((cap + w) / TB_STEP + 3) * TB_K3C_FREC * sizeof(double);
    y.total = (int)o;
    return y;
}

// FAST = 2: all sums O(1); FAST = 1: the rolling sum of the bias prediction keeps the literal float accumulator (it is the denominator of
// `expected`, whose ~3e-7 scatter the subtraction uncorrected - expected can amplify), the rescaling sums are O(1); FAST = 0: all literal.
// float32 strand sums of the four tracks, written next to the fp64 per-strand tracks when a sink asks for them (NULL: track not wanted):
// the reverse-strand pass of a tile adds the forward value the same thread stored before -- what tb_pack_track_kernel would compute later
struct tb_k3c_pk {
    float *p[4];
};

template <int FAST>
__global__ void __launch_bounds__(TB_K3C_THREADS, 3)
    tb_correct_kernel(tb_genome_view g, tb_cor_view V, const int32_t *__restrict__ tile_region,
                      const int32_t *__restrict__ tile_start, i64 n_tiles, int tile_len, const u32 *__restrict__ pile,
                      const double *__restrict__ biasraw, const double *__restrict__ fit, double cf, tb_k3c_layout Y,
                      double *__restrict__ tracks, double *__restrict__ prepost, tb_k3c_pk PK) {
    TB_DYN_SMEM(smem);
    const int w = V.window, k = V.k_flank, L = V.L;
    const int lf = w / 2;
    const int overlaps = w / TB_STEP;
    const int cap = tile_len + 4 * w;
    double *sig = (double *)smem;              // [d0, d1)  signal;            phase 4: pre weights of [t0, t1).  FAST: prefix sums of bp, then of corrected+
    double *bp = sig + Y.capd;                 // [d0, d1)  mean prediction;   phase 4: post(-) weights, padded by L zeros
    double *unc = bp + Y.capd;                 // [c0, c1)  signal * cf;       phase 4: positions by base.  FAST: prefix sums of |corrected|
    double *cor = unc + Y.capd;                // [c0, c1)  uncorrected - expected
    double *expd = cor + Y.capd;               // [t0, t1)  expected;          phase 4: post(+) weights
    double *ver = (double *)(smem + Y.off_ver);        // [2 strands][3][4][L] verification accumulators of this block
    u32 *psum = (u32 *)(smem + Y.off_psum);            // [cap + 1] prefix sums of the signal over [d0, d1)
    u32 *pcnt = (u32 *)(smem + Y.off_pcnt);            // [cap + 1] prefix counts of its non-zero positions (exact form)
    u32 *wsum = (u32 *)(smem + Y.off_wsum);            // [2][32] warp totals of the block scan
    unsigned short *nzpos = (unsigned short *)(smem + Y.off_nzpos);     // [cap] non-zero positions (relative to d0; exact form)
    double *wpre = FAST ? (double *)(smem + Y.off_wpre) : sig;          // [tile_len] pre weights of phase 4
    uint8_t *bases = (uint8_t *)(smem + Y.off_bases);                   // [tile_len + L] base codes of [t0 - k, t1 + k)
    __shared__ int bcnt[4];                                     // positions per base in the tile (+- k_flank)
    __shared__ double rcnt[512 / TB_STEP + 2];                  // 1 / n for the row counts of the mean (n <= window / 10)
    __shared__ double wtot[64];                                 // FAST: warp totals of the fp64 block scans
    const int cap_w = (cap + w) / TB_STEP + 3;
    double *frec = (double *)(smem + Y.off_frec);               // [cap_w][6] mode, max(prediction), a, b, 1 / max of the windows over [d0, d1)
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;

    for (int i = tid; i < 2 * 3 * 4 * L; i += nt) ver[i] = 0.0;
    for (int i = tid; i <= overlaps; i += nt) rcnt[i] = i > 0 ? 1.0 / (double)i : 0.0;
    __syncthreads();

    for (i64 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const i64 j = tile_region[tile];
        const int t0 = tile_start[tile];
        const int Lr = (int)(V.ext_off[j + 1] - V.ext_off[j]);
        const int nwin = tb_n_windows(Lr, k, w);
        const int t1 = min(Lr, t0 + tile_len);                    // tile = [t0, t1)
        const int c0 = max(0, t0 - w), c1 = min(Lr, t1 + w);      // domain of the raw corrected signal
        const int d0 = max(0, c0 - w), d1 = min(Lr, c1 + w);      // domain of signal / bias prediction
        const int nd = d1 - d0;
        const i64 gext = V.ext_off[j];
        const i64 gstart = V.ext_start[j];
        // base codes of the tile (+- k_flank), shared by both strands
        for (int i = tid; i < t1 - t0 + 2 * k; i += nt) {
            i64 gp = gstart + t0 - k + i;
            bases[i] = (gp >= 0 && gp < g.n_bases) ? (uint8_t)tb_base(g, gp) : (uint8_t)0;
        }

        for (int strand = 0; strand < 2; strand++) {
            const u32 *ps = pile + (i64)strand * V.ext_total + gext;
            const double *br = biasraw + (i64)strand * V.ext_total + gext;
            const double *ft = fit + ((i64)strand * V.n_windows + V.win_off[j]) * TB_FIT_REC;
            // ---- 1. signal, bias_prediction = nanmean over the `overlaps` rows (:279-309), prefix sums
            const int per = (nd + nt - 1) / nt;                   // consecutive positions per thread in the scan
            const int jw_lo = max(0, (d0 - k - w) / TB_STEP);     // windows that can cover a position of [d0, d1)
            {
                const int jw_hi = min(nwin - 1, max(0, d1 - 1 - k) / TB_STEP);
                const int nrec = min(cap_w, jw_hi - jw_lo + 1);
                for (int i = tid; i < nrec * 4; i += nt) {                           // read after the scan's barrier
                    const double v = ft[(i64)(jw_lo + (i >> 2)) * TB_FIT_REC + (i & 3)];         // mode, a, b, max(prediction)
                    double *fr = frec + (i >> 2) * TB_K3C_FREC;
                    const int c = i & 3;
                    fr[c == 0 ? 0 : (c == 3 ? 1 : c + 1)] = v;                       // -> mode, max, a, b
                    if (c == 3) fr[4] = v > 0.0 ? 1.0 / v : 0.0;                     // one division per window instead of one per covered position
                }
            }
            {
                u32 s_loc = 0, c_loc = 0;
                for (int e = 0; e < per; e++) {
                    int i = tid * per + e;
                    if (i < nd) {
                        u32 v = ps[d0 + i];
                        s_loc += v;
                        c_loc += v != 0u;
                    }
                }
                u32 s_inc = s_loc, c_inc = c_loc;
                for (int d = 1; d < 32; d <<= 1) {
                    u32 a = __shfl_up_sync(0xffffffffu, s_inc, d), b = __shfl_up_sync(0xffffffffu, c_inc, d);
                    if (lane >= d) {
                        s_inc += a;
                        c_inc += b;
                    }
                }
                if (lane == 31) {
                    wsum[warp] = s_inc;
                    wsum[32 + warp] = c_inc;
                }
                __syncthreads();
                u32 s_off = s_inc - s_loc, c_off = c_inc - c_loc;
                for (int ww = 0; ww < warp; ww++) {
                    s_off += wsum[ww];
                    c_off += wsum[32 + ww];
                }
                if (tid == 0) {
                    psum[0] = 0;
                    if (!FAST) pcnt[0] = 0;
                }
                for (int e = 0; e < per; e++) {
                    int i = tid * per + e;
                    if (i < nd) {
                        u32 v = ps[d0 + i];
                        s_off += v;
                        psum[i + 1] = s_off;
                        if (!FAST) {
                            if (v != 0u) nzpos[c_off] = (unsigned short)i;
                            c_off += v != 0u;
                            pcnt[i + 1] = c_off;
                            sig[i] = (double)v;
                        }
                    }
                }
            }
            for (int q = d0 + tid; q < d1; q += nt) {
                double x = br[q];
                double sum = 0.0;
                int cnt = 0;
                int jmax = q >= k ? (q - k) / TB_STEP : -1;
                if (jmax > nwin - 1) jmax = nwin - 1;
                const int r0 = jmax >= 0 ? jmax % overlaps : 0;
                for (int r = 0; r < overlaps; r++) {
                    double val = 0.0;
                    bool have = false;
                    if (jmax >= r) {
                        const int jw = jmax - r0 + r - (r > r0 ? overlaps : 0);         // last window of row r starting at or before q
                        if (k + TB_STEP * jw + w > q) {
                            have = true;
                            const double *f = frec + (jw - jw_lo) * TB_K3C_FREC;
                            const double mode = f[0], mx = f[1], minv = f[4];
                            if (mode == 0.0) {
                                double t = __dadd_rn(__dmul_rn(f[2], x), f[3]);
                                val = t > 0.0 ? t : 0.0;
                                if (mx > 0.0) val = tb_div_inv(val, mx, minv);      // = val / mx, correctly rounded (tb_lm_thread.cuh)
                            } else if (mode == 1.0) {
                                double t = x - f[3];
                                val = t < 0.0 ? 0.0 : t;
                                if (mx > 0.0) val = tb_div_inv(val, mx, minv);
                            }
                        }
                    }
                    // rows k_flank.. start as NaN (:280 indexes rows): uncovered cells there do not count
                    if (have || r < k) {
                        sum += val;
                        cnt++;
                    }
                }
                const double m = cnt > 0 ? tb_div_inv(sum, (double)cnt, rcnt[cnt]) : (double)NAN;  // = sum / cnt
                bp[q - d0] = m;
                if (FAST == 2) {
                    sig[q - d0] = m;
                    unc[q - d0] = 0.0;
                }
            }
            __syncthreads();
            if (FAST == 2) tb_block_scan2w(sig, unc, nd, wtot);    // sig := prefix sums of the bias prediction (unc rides along, rewritten below)
            // ---- 2. expected and raw corrected signal on [c0, c1) (:314-328)
            for (int q = c0 + tid; q < c1; q += nt) {
                double ssum = 0.0, bsum = 1.0, probas = 0.0;
                if (q >= lf && q - lf + w <= Lr) {
                    const int first = q - lf - d0;
                    const u32 isum = psum[first + w] - psum[first];
                    double b;
                    if (FAST == 2) {
                        ssum = isum < (1u << 24) ? (double)isum : tb_round_f32((double)isum);
                        b = tb_round_f32(sig[first + w] - sig[first]);
                    } else if (FAST == 1) {
                        ssum = isum < (1u << 24) ? (double)isum : tb_round_f32((double)isum);
                        b = tb_roll_f32(bp, first, w);
                    } else {
                        ssum = isum < (1u << 24) ? (double)isum : tb_roll_f32(sig, first, w);
                        b = tb_roll_f32(bp, first, w);
                    }
                    bool null = !(fabs(b) > 1e-8);                 // isclose(bias_sum, 0) or NaN
                    if (!null) {
                        bsum = b;
                        probas = bp[q - d0] / bsum;
                    }
                }
                double e = ssum * probas;
                double u = (FAST ? (double)(psum[q - d0 + 1] - psum[q - d0]) : sig[q - d0]) * cf;
                e = e * cf;
                if (!FAST) unc[q - c0] = u;
                cor[q - c0] = u - e;
                if (q >= t0 && q < t1) expd[q - t0] = e;
            }
            __syncthreads();
            if (FAST) {      // prefix sums of |corrected| (in unc) and of its positive part (in sig: the prefix sums of bp are no longer read)
                const int nc = c1 - c0;
                for (int i = tid; i < nc; i += nt) {
                    const double cv = cor[i];
                    unc[i] = fabs(cv);
                    sig[i] = cv > 0.0 ? cv : 0.0;
                }
                __syncthreads();
                tb_block_scan2w(unc, sig, nc, wtot);
            }
            // ---- 3. rescale (:331-350), outputs, verification weights (:358-373)
            for (int q = t0 + tid; q < t1; q += nt) {
                double c = cor[q - c0];
                double u = FAST ? (double)(psum[q - d0 + 1] - psum[q - d0]) * cf : unc[q - c0];
                double scale = 1.0;
                if (q >= lf && q - lf + w <= Lr) {
                    int first = q - lf - c0;
                    double au = 0.0, aa = 0.0, ap = 0.0;            // float-valued accumulators (tb_round_f32)
                    if (FAST) {
                        const int fd = q - lf - d0;
                        aa = tb_round_f32(unc[first + w] - unc[first]);
                        ap = tb_round_f32(sig[first + w] - sig[first]);
                        au = tb_round_f32((double)(psum[fd + w] - psum[fd]) * cf);
                    } else {
                        for (int t = 0; t < w; t++) {
                            double cv = cor[first + t];
                            aa = tb_round_f32(__dadd_rn(aa, fabs(cv)));
#ifdef TB_K3C_MIX                                       // comparison build: this sum through the conversion unit, the other in the fp64 pipe
                            if (!(cv <= 0.0)) ap = (double)(float)__dadd_rn(ap, cv);
#else
                            if (!(cv <= 0.0)) ap = tb_round_f32(__dadd_rn(ap, cv));     // the non-positive half adds 0.0: no change
#endif
                        }
                        {   // uncorrected: only non-zero signal positions move the accumulator
                            const int fd = q - lf - d0;
                            for (u32 z = pcnt[fd]; z < pcnt[fd + w]; z++) au = tb_round_f32(__dadd_rn(au, unc[(int)nzpos[z] + d0 - c0]));
                        }
                    }
                    double usum = au, asum = aa, psum_ = ap;
                    double nsum = asum - psum_;
                    if (psum_ != 0.0) {
                        scale = (usum - nsum) / psum_;
                        if (scale < 1.0) scale = 1.0;
                    }
                }
                if (c > 0.0) c = c * scale;
                const double e = expd[q - t0];
                if (q >= V.f_extend && q < Lr - V.f_extend) {
                    i64 o = V.out_off[j] + (q - V.f_extend);
                    tracks[((i64)0 * 2 + strand) * V.out_total + o] = u;
                    tracks[((i64)1 * 2 + strand) * V.out_total + o] = br[q];
                    tracks[((i64)2 * 2 + strand) * V.out_total + o] = e;
                    tracks[((i64)3 * 2 + strand) * V.out_total + o] = c;
                    if (strand == 1) {
                        if (PK.p[0]) PK.p[0][o] = (float)(tracks[((i64)0 * 2) * V.out_total + o] + u);
                        if (PK.p[1]) PK.p[1][o] = (float)(tracks[((i64)1 * 2) * V.out_total + o] + br[q]);
                        if (PK.p[2]) PK.p[2][o] = (float)(tracks[((i64)2 * 2) * V.out_total + o] + e);
                        if (PK.p[3]) PK.p[3][o] = (float)(tracks[((i64)3 * 2) * V.out_total + o] + c);
                    }
                }
                bool valid = q > k && q < Lr - k - 1 && (u != 0.0 || c != 0.0);
                if (valid) valid = tb_nflags(g, gstart + q - k, L) == 0;
                // wpre (exact form: sig) / bp are not read in this phase: they now carry the weights of phase 4
                wpre[q - t0] = (valid && u > 0.0) ? u : 0.0;              // pre_bias counts
                expd[q - t0] = (valid && c > 0.0) ? c : 0.0;              // post_bias counts
                bp[L + q - t0] = (valid && !(c > 0.0)) ? c : 0.0;         // post_bias neg_counts, L zeros on either side
            }
            {
                const int n_pos = t1 - t0;
                if (tid < L) bp[tid] = 0.0;
                else if (tid < 2 * L) bp[n_pos + tid] = 0.0;
                // positions of [t0 - k, t1 + k) by base, increasing (in the memory of `unc`, which the phase above does not read ... but
                // other threads may still be in it: the lists are written after the barrier)
            }
            __syncthreads();
            unsigned short *blist = (unsigned short *)unc;               // [4][tile_len + L]
            const int bl_cap = tile_len + L;
            if (warp < 4) {
                const int nb = t1 - t0 + 2 * k;
                int cnt = 0;
                for (int i = 0; i < nb; i += 32) {
                    const bool mine = i + lane < nb && bases[i + lane] == warp;
                    const unsigned m = __ballot_sync(0xffffffffu, mine);
                    if (mine) blist[warp * bl_cap + cnt + __popc(m & ((1u << lane) - 1u))] = (unsigned short)(i + lane);
                    cnt += __popc(m);
                }
                if (lane == 0) bcnt[warp] = cnt;
            }
            __syncthreads();
            // ---- 4. verification counts: lane m owns offset m of the k-mer.  pre / post(+) weights are sparse: the positions of the warp's
            //         strip that carry one are broadcast and added under the base of the lane's k-mer position.  post(-) weights are dense
            //         (every covered position without reads): per base, the lane walks the list of positions holding that base and adds the
            //         weight of the k-mer centre that puts its offset there -- one load and one add per (position, offset), no predication
            {
                const int n_pos = t1 - t0;
                const int strip = (n_pos + nwarps - 1) / nwarps;
                const int qa = warp * strip, qb = min(n_pos, qa + strip);
                double acc[3][4];
#pragma unroll
                for (int a = 0; a < 3; a++)
#pragma unroll
                    for (int s = 0; s < 4; s++) acc[a][s] = 0.0;
                // the sparse weights: 32 positions of the strip are tested per step, the warp visits the ones that carry a weight (in order)
                for (int q0 = qa; q0 < qb; q0 += 32) {
                    const int qq = q0 + lane;
                    const bool nz = qq < qb && (wpre[qq] != 0.0 || expd[qq] != 0.0);
                    unsigned hit = __ballot_sync(0xffffffffu, nz);
                    while (hit) {
                        const int q = q0 + __ffs(hit) - 1;
                        hit &= hit - 1;
                        if (lane < L) {
                            const double v0 = wpre[q], v1 = expd[q];
                            const int s = bases[q + lane];
#pragma unroll
                            for (int b = 0; b < 4; b++) {
                                if (s == b) {
                                    acc[0][b] += v0;                                  // adding 0.0 changes nothing
                                    acc[1][b] += v1;
                                }
                            }
                        }
                    }
                }
                if (lane < L) {
                    const double *negw = bp + L - lane;                               // negw[p] = weight of centre p - lane (0 outside the tile)
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        const int nb = bcnt[b];
                        const int per = (nb + nwarps - 1) / nwarps;
                        const int ia = warp * per, ib = min(nb, ia + per);
                        const unsigned short *lst = blist + b * bl_cap;
                        double a2 = 0.0;
#pragma unroll 4
                        for (int i = ia; i < ib; i++) a2 += negw[lst[i]];
                        acc[2][b] = a2;
                    }
                    const int mm = strand ? L - 1 - lane : lane;
#pragma unroll
                    for (int a = 0; a < 3; a++)
#pragma unroll
                        for (int b = 0; b < 4; b++)
                            if (acc[a][b] != 0.0) atomicAdd(&ver[((strand * 3 + a) * 4 + (strand ? (b ^ 1) : b)) * L + mm], acc[a][b]);
                }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < 2 * 3 * 4 * L; i += nt) {
        double v = ver[i];
        if (v != 0.0) {
            int sw = i / (4 * L), rem = i % (4 * L);          // sw = strand * 3 + which
            atomicAdd(&prepost[(i64)sw * 5 * L + rem], v);
        }
    }
}

// out[i] = tracks[t][0][i] (+ tracks[t][1][i]) as float or double
template <class T>
__global__ void tb_pack_track_kernel(const double *__restrict__ a, const double *__restrict__ b, i64 n, T *__restrict__ out) {
    i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (T)(b ? a[i] + b[i] : a[i]);
}

// ---------------------------------------------------------------------------------------------- host side
// Packs the output range [o0, o1) of the tracks the sink asks for and copies it to the host, on the copy stream, ordered after the work the
// compute stream holds so far.  The last chunk is followed by the verification counts.
static int tb_sink_chunk(tb_ctx *ctx, int chunk, i64 o0, i64 o1, bool last, bool packed) {
    auto &C = ctx->cor;
    const size_t esz = C.sink_f64 ? 8 : 4;
    const i64 n = C.out_total, len = o1 - o0;
    TB_CUDA(ctx, cudaEventRecord(ctx->ev_chunk[chunk & 7], ctx->stream));
    TB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_chunk[chunk & 7], 0));
    if (len > 0) {
        for (int t = 0; t < 4 && !packed; t++)
            for (int s = 0; s < 2; s++) {
                if (!C.sink_ptr[t][s]) continue;
                const double *a = C.tracks.as<double>() + ((i64)t * 2 + (C.sink_split ? s : 0)) * n + o0;
                const double *b = C.sink_split ? nullptr : C.tracks.as<double>() + ((i64)t * 2 + 1) * n + o0;
                i64 want = tb_div_up(len, 256), capg = (i64)ctx->sm_count * 4;
                unsigned grid = (unsigned)(want < capg ? want : capg);
                if (C.sink_f64) {
                    auto kp = tb_pack_track_kernel<double>;
                    TB_LAUNCH_ON(ctx, ctx->copy_stream, kp, grid, 256, 0, a, b, len, C.dl[t][s].as<double>() + o0);
                } else {
                    auto kp = tb_pack_track_kernel<float>;
                    TB_LAUNCH_ON(ctx, ctx->copy_stream, kp, grid, 256, 0, a, b, len, C.dl[t][s].as<float>() + o0);
                }
            }
        for (int t = 0; t < 4; t++)
            for (int s = 0; s < 2; s++)
                if (C.sink_ptr[t][s])
                    TB_CUDA(ctx, cudaMemcpyAsync((char *)C.sink_ptr[t][s] + (size_t)o0 * esz, (const char *)C.dl[t][s].p + (size_t)o0 * esz,
                                                 (size_t)len * esz, cudaMemcpyDeviceToHost, ctx->copy_stream));
    }
    if (last) {
        C.pp_dst = C.sink_pp;
        if (C.sink_pp) {
            if (!C.pp_pinned) TB_CUDA(ctx, cudaMallocHost((void **)&C.pp_pinned, (size_t)2 * 3 * 5 * 32 * sizeof(double)));
            TB_CUDA(ctx, cudaMemcpyAsync(C.pp_pinned, C.prepost.p, (size_t)2 * 3 * 5 * C.L * sizeof(double), cudaMemcpyDeviceToHost,
                                         ctx->copy_stream));
        }
    }
    return TB_OK;
}

extern "C" int tb_correct_set_sink(tb_ctx *ctx, int split_strands, int as_f64, void *tracks[4][2], double *prepost_out) {
    if (!ctx) return TB_ERR_ARG;
    auto &C = ctx->cor;
    C.sink_on = false;
    if (!tracks) return TB_OK;
    for (int t = 0; t < 4; t++)
        for (int s = 0; s < 2; s++) {
            if (tracks[t][s] && !split_strands && s == 1)
                return tb_fail(ctx, TB_ERR_ARG, "tb_correct_set_sink: tracks[t][1] must be NULL unless split_strands");
            C.sink_ptr[t][s] = tracks[t][s];
        }
    C.sink_split = split_strands ? 1 : 0;
    C.sink_f64 = as_f64 ? 1 : 0;
    C.sink_pp = prepost_out;
    C.sink_on = true;
    return TB_OK;
}

extern "C" int tb_correct_run(tb_ctx *ctx, int64_t n_regions, const int32_t *contig, const int64_t *start, const int64_t *end,
                              int k_flank, int window, double correction_factor, int shift_fwd, int shift_rev,
                              int lm_exact_order) {
    if (!ctx) return TB_ERR_ARG;
    const bool trace = tb_opt(ctx, "host_trace", 0) != 0;        // host clock at the phase borders of this call, to stderr
    const auto t_entry = std::chrono::steady_clock::now();
    auto tr = [&](const char *what) {
        if (trace) fprintf(stderr, "[tb_correct_run] %-28s %9.3f ms\n", what,
                           std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_entry).count());
    };
    ctx->cor.valid = false;
    TB_TRY(tb_sync(ctx, "tb_correct_run (entry)"));              // a kernel of an earlier call may still read the tables this run replaces
    const bool sink_now = ctx->cor.sink_on;                      // a sink serves this run only, whatever its outcome
    ctx->cor.sink_on = false;
    ctx->cor.pp_dst = nullptr;
    if (window < TB_STEP || window > 512) return tb_fail(ctx, TB_ERR_ARG, "tb_correct_run: window must be in [10, 512]");
    if (ctx->model[0].is_dwm < 0 || ctx->model[1].is_dwm < 0)
        return tb_fail(ctx, TB_ERR_STATE, "tb_correct_run: call tb_set_bias_model for both strands first");
    const int L = ctx->model[0].L;
    if (L != 2 * k_flank + 1) return tb_fail(ctx, TB_ERR_ARG, "tb_correct_run: k_flank %d does not match the model length %d", k_flank, L);
    std::vector<i64> &ls = ctx->cor.h_ls, &le = ctx->cor.h_le;
    TB_TRY(tb_linear_regions(ctx, n_regions, contig, start, end, ls, le, false));
    tr("linear regions");
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
    std::vector<i64> &ext_end = C.h_ext_end;
    std::vector<int32_t> &tile_region = C.h_tile_region, &tile_start = C.h_tile_start;
    ext_end.resize(n_regions);
    tile_region.clear();
    tile_start.clear();
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
    }
    tr("offsets");
    C.ext_total = C.ext_off[n_regions];
    C.out_total = C.out_off[n_regions];
    C.n_windows = C.win_off[n_regions];
    if (n_regions == 0) {
        C.valid = true;
        return TB_OK;
    }
    // the tables the first two kernels read go up now; the rest (output / window offsets, tiles of the last stage) is prepared on the host
    // while those kernels run and travels on the copy stream
    TB_TRY(tb_h2d(ctx, C.d_ext_start, C.ext_start.data(), sizeof(i64) * n_regions));
    TB_TRY(tb_h2d(ctx, C.d_ext_end, ext_end.data(), sizeof(i64) * n_regions));
    TB_TRY(tb_h2d(ctx, C.d_ext_off, C.ext_off.data(), sizeof(i64) * (n_regions + 1)));
    TB_TRY(tb_reserve(ctx, C.pile, (size_t)C.ext_total * 2 * sizeof(u32)));
    TB_TRY(tb_reserve(ctx, C.biasraw, (size_t)C.ext_total * 2 * sizeof(double)));
    TB_TRY(tb_reserve(ctx, C.fit, (size_t)C.n_windows * 2 * TB_FIT_REC * sizeof(double)));
    TB_TRY(tb_reserve(ctx, C.tracks, (size_t)C.out_total * 8 * sizeof(double)));
    TB_TRY(tb_reserve(ctx, C.prepost, (size_t)2 * 3 * 5 * L * sizeof(double)));
    TB_TRY(tb_reserve(ctx, C.d_out_off, sizeof(i64) * (n_regions + 1)));
    TB_TRY(tb_reserve(ctx, C.d_win_off, sizeof(i64) * (n_regions + 1)));
    TB_CUDA(ctx, cudaMemsetAsync(C.prepost.p, 0, (size_t)2 * 3 * 5 * L * sizeof(double), ctx->stream));
    if (sink_now)
        for (int t = 0; t < 4; t++)
            for (int s = 0; s < 2; s++)
                if (C.sink_ptr[t][s]) TB_TRY(tb_reserve(ctx, C.dl[t][s], (size_t)C.out_total * (C.sink_f64 ? 8 : 4)));

    tr("tables uploaded, buffers");
    tb_timer_start(ctx);
    tb_mark(ctx, 0);
    TB_TRY(tb_pileup_device(ctx, n_regions, C.d_ext_start.as<i64>(), C.d_ext_end.as<i64>(), C.d_ext_off.as<i64>(), C.ext_total,
                            shift_fwd, shift_rev, C.pile.as<u32>()));
    tb_mark(ctx, 1);
    TB_TRY(tb_score_device(ctx, C.d_ext_start.as<i64>(), C.d_ext_off.as<i64>(), n_regions, C.ext_total, 3, 0,
                           C.biasraw.as<double>(), C.biasraw.as<double>() + C.ext_total));
    tb_mark(ctx, 2);
    tr("K1 + K3a enqueued");
    // K3c tile length: every tile recomputes a halo of 2 * window positions on each side, so the tile that covers most regions in
    // one piece wins (hg38 peaks: 824-bp regions -> one 832-position tile each instead of two 512-position tiles with halos)
    int tile_len = 512;
    {
        std::vector<std::pair<i64, i64>> hist;               // runs of equal region lengths
        for (i64 i = 0; i < n_regions; i++) {
            const i64 v = C.ext_off[i + 1] - C.ext_off[i];
            if (!hist.empty() && hist.back().first == v) hist.back().second++;
            else hist.push_back({v, 1});
        }
        double best = -1.0;
        for (int cand = 256; cand <= TB_K3C_TILE_MAX; cand += 64) {
            double cost = 0.0;
            for (const auto &h : hist) {
                const i64 nt = tb_div_up(h.first, cand);
                const i64 halo = nt > 1 ? 4 * (i64)window * (nt - 1) : 0;       // positions computed twice for the interior borders
                cost += (double)(h.first + halo + 32 * nt) * (double)h.second;  // + a per-tile constant (barriers, scans)
            }
            cost *= 1.0 + 0.15 * (double)cand / TB_K3C_TILE_MAX;                // larger tiles: fewer resident blocks
            if (best < 0.0 || cost < best) {
                best = cost;
                tile_len = cand;
            }
        }
    }
    tr("tile length chosen");
    {
        i64 est = 0;
        for (i64 i = 0; i < n_regions; i++) est += tb_div_up(C.ext_off[i + 1] - C.ext_off[i], tile_len);
        tile_region.reserve((size_t)est);
        tile_start.reserve((size_t)est);
    }
    for (i64 i = 0; i < n_regions; i++) {
        const i64 reg_len = C.ext_off[i + 1] - C.ext_off[i];
        for (i64 t = 0; t < reg_len; t += tile_len) {
            tile_region.push_back((int32_t)i);
            tile_start.push_back((int32_t)t);
        }
    }
    tr("host tables built");
    TB_TRY(tb_h2d_aux(ctx, C.d_out_off, C.out_off.data(), sizeof(i64) * (n_regions + 1)));
    TB_TRY(tb_h2d_aux(ctx, C.d_win_off, C.win_off.data(), sizeof(i64) * (n_regions + 1)));
    TB_TRY(tb_h2d_aux(ctx, ctx->reg_a, tile_region.data(), sizeof(int32_t) * tile_region.size()));
    TB_TRY(tb_h2d_aux(ctx, ctx->reg_b, tile_start.data(), sizeof(int32_t) * tile_start.size()));
    TB_TRY(tb_aux_join(ctx));
    tr("second tables uploaded");
    tb_cor_view V{C.d_ext_start.as<i64>(), C.d_ext_off.as<i64>(), C.d_out_off.as<i64>(), C.d_win_off.as<i64>(),
                  n_regions, C.ext_total, C.out_total, C.n_windows, k_flank, window, L, f_extend};
    // window fits: lm_exact_order 0 -> one-pass fits with forward-difference inner products (fit_form 1: the DWM default), 1 -> MINPACK-exact
    // (fit_form 0), 5 -> moment form of the one-pass fits (fit_form 2: exact derivative instead of difference quotients -- a third of the
    // arithmetic, but it leaves scipy's path on more of the windows that scipy itself does not fit stably, and the round kernels are bound
    // by the L2 traffic of re-reading every window once per pass, not by arithmetic: measured 52.1 vs 53.6 ms; opt-in); 2 / 3:
    // first-generation warp kernels (4 is accepted as an alias of 0)
    const int fit_form = lm_exact_order == 5 ? 2 : ((lm_exact_order == 0 || lm_exact_order == 4) ? 1 : 0);
    if (lm_exact_order < 2 || lm_exact_order == 4 || lm_exact_order == 5) {
        // round form (default): state in HBM, compacted work lists, one host read-back of two counters per round
        const int qm = (window + 9) / 10;
        const i64 qs = C.n_windows + (i64)(qm - 1) * n_regions, n_slots = 2 * C.n_windows;
        // persistent form (one-pass fits, window = 100): tiles of TS slots in shared memory; the planes are padded so that every tile (and the 12
        // slots behind it) can be read with aligned 16-byte loads
        // (persistent form: forward-difference one-pass fits only; measured slower than the rounds on hg38 -- 75.6 vs 53.6 ms -- and off by default)
        int persistent = (fit_form == 1 && qm == 10) ? (int)tb_opt(ctx, "k3b_persistent", 0) : 0;
        const int fitp_threads = tb_opt(ctx, "k3b_threads", 512) == 256 ? 256 : 512;      // 512: one block per SM; 256: two, half the tile each
#ifdef TB_EMU
        const int fitp_ts_default = 256;                          // the test emulator runs small tiles (same code)
#else
        const int fitp_ts_default = fitp_threads == 256 ? 800 : 1600;
#endif
        const int fitp_ts = tb_opt(ctx, "k3b_tile", 0) >= 64 ? (int)(tb_opt(ctx, "k3b_tile", 0) & ~31LL) : fitp_ts_default;
        const i64 n_ptiles = tb_div_up(2 * qs, fitp_ts);
        const i64 qtot = persistent ? n_ptiles * fitp_ts + 16 : 2 * qs;
        if (10 * qtot >= ((i64)1 << 32)) return tb_fail(ctx, TB_ERR_ARG, "tb_correct_run: too many windows for one launch");
        TB_TRY(tb_reserve(ctx, C.fitg_x, (size_t)10 * qtot * sizeof(double)));
        TB_TRY(tb_reserve(ctx, C.fitg_y, (size_t)10 * qtot * sizeof(u32)));
        if (persistent) {
            TB_TRY(tb_reserve(ctx, C.fitg_sw, (size_t)qtot * sizeof(int32_t)));
            TB_CUDA(ctx, cudaMemsetAsync(C.fitg_sw.p, 0xff, (size_t)qtot * sizeof(int32_t), ctx->stream));      // -1: no window
        }
        TB_TRY(tb_reserve(ctx, C.fitg_st, (size_t)TB_FITG_NST * n_slots * sizeof(double)));
        TB_TRY(tb_reserve(ctx, C.fitg_sti, (size_t)n_slots * sizeof(int4)));
        TB_TRY(tb_reserve(ctx, C.fitg_wq, (size_t)n_slots * sizeof(u32)));
        TB_TRY(tb_reserve(ctx, C.fitg_lists, (size_t)4 * n_slots * sizeof(u32) + 64));
        u32 *lists = C.fitg_lists.as<u32>();
        u32 *lq[2] = {lists, lists + n_slots}, *lt[2] = {lists + 2 * n_slots, lists + 3 * n_slots};
        unsigned long long *ctr = (unsigned long long *)(lists + 4 * n_slots);       // [0] Q, [1] T of the list being built
        tb_fitg_view G{C.fitg_st.as<double>(), C.fitg_sti.as<int4>(), C.fitg_wq.as<u32>(), C.fitg_x.as<double>(),
                       C.fitg_y.as<u32>(), qtot, n_slots};
        TB_LAUNCH(ctx, tb_fitg_transpose_kernel, dim3((unsigned)n_regions, 2), 256, 0, V, (const u32 *)C.pile.as<u32>(),
                  (const double *)C.biasraw.as<double>(), qs, qtot, qm, C.fitg_x.as<double>(), C.fitg_y.as<u32>(),
                  persistent ? C.fitg_sw.as<int32_t>() : (int32_t *)nullptr);
        TB_CUDA(ctx, cudaMemsetAsync(ctr, 0, 2 * sizeof(unsigned long long), ctx->stream));
        if (persistent) {
            const tb_fitp_tile T{fitp_ts, fitp_ts + 12};
            const size_t psmem = (size_t)10 * T.tsp * (sizeof(double) + sizeof(u32));
            const int cap_passes = tb_opt(ctx, "k3b_cap", 0) > 0 ? (int)tb_opt(ctx, "k3b_cap", 0) : 8;
            if (psmem > (size_t)226 * 1024) return tb_fail(ctx, TB_ERR_ARG, "tb_correct_run: k3b_tile too large for the shared memory");
            if (fitp_threads == 256) {
                auto kp = tb_fitp_kernel<256, 2>;
                TB_CUDA(ctx, cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
                const i64 capg = (i64)ctx->sm_count * (psmem > (size_t)112 * 1024 ? 1 : 2);
                TB_LAUNCH(ctx, kp, (unsigned)(n_ptiles < capg ? n_ptiles : capg), 256, psmem, G, (const int32_t *)C.fitg_sw.as<int32_t>(), T,
                          n_ptiles, window, cap_passes, C.fit.as<double>(), lt[0], ctr);
            } else {
                auto kp = tb_fitp_kernel<512, 1>;
                TB_CUDA(ctx, cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
                const i64 capg = (i64)ctx->sm_count * (psmem > (size_t)112 * 1024 ? 1 : 2);
                TB_LAUNCH(ctx, kp, (unsigned)(n_ptiles < capg ? n_ptiles : capg), 512, psmem, G, (const int32_t *)C.fitg_sw.as<int32_t>(), T,
                          n_ptiles, window, cap_passes, C.fit.as<double>(), lt[0], ctr);
            }
        } else {
            i64 want = tb_div_up(n_slots, 128), capg = (i64)ctx->sm_count * 16;
            if (fit_form == 0) {
                TB_LAUNCH(ctx, tb_fitg_init_kernel<0>, (unsigned)(want < capg ? want : capg), 128, 0, V, G, qs, qm, C.fit.as<double>(), lq[0], ctr);
            } else if (fit_form == 1) {      // one-pass forms: the first head ran in the init pass, every window starts on the trial-only list
                TB_LAUNCH(ctx, tb_fitg_init_kernel<1>, (unsigned)(want < capg ? want : capg), 128, 0, V, G, qs, qm, C.fit.as<double>(), lt[0], ctr + 1);
            } else {
                TB_LAUNCH(ctx, tb_fitg_init_kernel<2>, (unsigned)(want < capg ? want : capg), 128, 0, V, G, qs, qm, C.fit.as<double>(), lt[0], ctr + 1);
            }
        }
        unsigned long long hc[2] = {0, 0};
        TB_TRY(tb_d2h(ctx, hc, ctr, sizeof(hc)));
        i64 nq = (i64)hc[0], nt = (i64)hc[1];
        int cur = 0;
        const i64 tail_at = (i64)ctx->sm_count * 128;      // one block per SM's worth of windows: finish them in one launch
        for (int round = 0; nq + nt > 0; round++) {
            if (round > 700) return tb_fail(ctx, TB_ERR_STATE, "tb_correct_run: window fits did not terminate");
            const i64 capg = (i64)ctx->sm_count * (fit_form == 0 ? TB_FITG_MINBLOCKS : TB_FITG_MINBLOCKS + 1);      // resident blocks (launch bounds of the round kernels)
            if (nq + nt <= tail_at) {
                i64 want = tb_div_up(nq + nt, 128);
                auto kf = fit_form == 0 ? tb_fitg_finish_kernel<0> : (fit_form == 1 ? tb_fitg_finish_kernel<1> : tb_fitg_finish_kernel<2>);
                TB_LAUNCH(ctx, kf, (unsigned)(want < capg ? want : capg), 128, 0, G, window, (const u32 *)lq[cur], nq, (const u32 *)lt[cur], nt,
                          C.fit.as<double>());
                break;
            }
            TB_CUDA(ctx, cudaMemsetAsync(ctr, 0, 2 * sizeof(unsigned long long), ctx->stream));
            if (nq > 0) {
                i64 want = tb_div_up(nq, 128);
                auto kq = fit_form == 0 ? tb_fitg_round_kernel<true, 0> : (fit_form == 1 ? tb_fitg_round_kernel<true, 1> : tb_fitg_round_kernel<true, 2>);
                TB_LAUNCH(ctx, kq, (unsigned)(want < capg ? want : capg), 128, 0, G, window, (const u32 *)lq[cur], nq, C.fit.as<double>(),
                          lq[cur ^ 1], lt[cur ^ 1], ctr);
            }
            if (nt > 0) {
                i64 want = tb_div_up(nt, 128);
                auto kt = fit_form == 0 ? tb_fitg_round_kernel<false, 0> : (fit_form == 1 ? tb_fitg_round_kernel<false, 1> : tb_fitg_round_kernel<false, 2>);
                TB_LAUNCH(ctx, kt, (unsigned)(want < capg ? want : capg), 128, 0, G, window, (const u32 *)lt[cur], nt, C.fit.as<double>(),
                          lq[cur ^ 1], lt[cur ^ 1], ctr);
            }
            TB_TRY(tb_d2h(ctx, hc, ctr, sizeof(hc)));
            nq = (i64)hc[0];
            nt = (i64)hc[1];
            cur ^= 1;
        }
        TB_TRY(tb_check(ctx, cudaGetLastError(), "tb_fitg_round_kernel"));
    } else {
#ifdef TB_FIRST_GEN_FITS
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
#else
        return tb_fail(ctx, TB_ERR_ARG, "tb_correct_run: lm_exact_order %d (first-generation warp-per-window fits) is not part of this build "
                       "(compile with -DTB_FIRST_GEN_FITS); use 0, 1 or 5", lm_exact_order);
#endif
    }
    tb_mark(ctx, 3);
    tr("K3b rounds done (host)");
    {
        TB_TRY(tb_genome_fence(ctx, 0, ctx->n_contigs - 1));
        tb_genome_view g{ctx->seq2.as<u32>(), ctx->nmask.as<u32>(), ctx->genome_bases};
        // the O(1) rolling sums need every row of the 10-row mean to count everywhere (k_flank >= window / 10: no NaN in the bias prediction)
        int fast = (fit_form != 0 && k_flank >= window / TB_STEP) ? 2 : 0;
        if (fast) fast = (int)tb_opt(ctx, "k3c_fast", fast);      // comparison runs: 0 literal, 1 literal bias-prediction sum only
        const tb_k3c_layout Y = tb_k3c_make_layout(tile_len, window, L, fast != 0);
        const size_t smem = (size_t)Y.total;
        auto kc = fast == 2 ? tb_correct_kernel<2> : (fast == 1 ? tb_correct_kernel<1> : tb_correct_kernel<0>);
        TB_CUDA(ctx, cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const i64 n_tiles = (i64)tile_region.size();
        i64 per_sm = (i64)(227 * 1024) / (i64)(smem + 1024);            // 227 KB usable per SM, 1 KB reserved per block
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 8) per_sm = 8;
        const i64 capg = (i64)ctx->sm_count * per_sm;
        // With a sink (tb_correct_set_sink) the tiles go in chunks that end on region borders; every chunk's tracks are packed and copied to
        // the host on the copy stream while the next chunk computes.
        const int n_chunks = sink_now ? (int)std::max<i64>(1, std::min<i64>(tb_opt(ctx, "k3c_chunks", 6), std::min<i64>(8, n_regions))) : 1;
        // strands combined as float32: the kernel writes the packed values itself, the chunk only has to be copied
        const bool fused_pack = sink_now && !C.sink_split && !C.sink_f64;
        tb_k3c_pk PK{{nullptr, nullptr, nullptr, nullptr}};
        if (fused_pack)
            for (int t = 0; t < 4; t++) PK.p[t] = C.sink_ptr[t][0] ? C.dl[t][0].as<float>() : nullptr;
        i64 t_done = 0, r_done = 0;
        for (int ch = 0; ch < n_chunks; ch++) {
            i64 r_end = ch + 1 == n_chunks ? n_regions : (n_regions * (ch + 1)) / n_chunks;
            if (r_end <= r_done) continue;
            i64 t_end = t_done;
            while (t_end < n_tiles && tile_region[t_end] < r_end) t_end++;
            const i64 cnt = t_end - t_done;
            if (cnt > 0)
                TB_LAUNCH(ctx, kc, (unsigned)(cnt < capg ? cnt : capg), TB_K3C_THREADS, smem, g, V, (const int32_t *)ctx->reg_a.as<int32_t>() + t_done,
                          (const int32_t *)ctx->reg_b.as<int32_t>() + t_done, cnt, tile_len, (const u32 *)C.pile.as<u32>(),
                          (const double *)C.biasraw.as<double>(), (const double *)C.fit.as<double>(), correction_factor, Y, C.tracks.as<double>(),
                          C.prepost.as<double>(), PK);
            if (sink_now) TB_TRY(tb_sink_chunk(ctx, ch, C.out_off[r_done], C.out_off[r_end], ch + 1 == n_chunks, fused_pack));
            t_done = t_end;
            r_done = r_end;
        }
        TB_TRY(tb_check(ctx, cudaGetLastError(), "tb_correct_kernel"));
    }
    tb_mark(ctx, 4);
    tr("K3c enqueued");
    tb_timer_stop(ctx);
    tr("timer stop (GPU drained)");
    tb_marks_collect(ctx, 5);      // stage_ms: [0] K1 pileup, [1] K3a score, [2] K3b fit, [3] K3c correct
    TB_TRY(tb_sync(ctx, "tb_correct_run"));
    tr("return");
    C.valid = true;
    return TB_OK;
}

// Enqueues the packing kernels and device-to-host copies of the requested tracks on the context's copy stream, ordered
// after everything already submitted to the compute stream.  Kernels launched afterwards (footprint scoring) overlap
// with the copies; tb_correct_download_wait blocks until the host buffers are complete.
extern "C" int tb_correct_download_begin(tb_ctx *ctx, int split_strands, int as_f64, void *tracks[4][2], double *prepost_out) {
    if (!ctx || !tracks) return TB_ERR_ARG;
    auto &C = ctx->cor;
    if (!C.valid) return tb_fail(ctx, TB_ERR_STATE, "tb_correct_download: no completed tb_correct_run");
    const size_t esz = as_f64 ? 8 : 4;
    const i64 n = C.out_total;
    for (int t = 0; t < 4; t++)
        for (int s = 0; s < 2; s++)
            if (tracks[t][s] && !split_strands && s == 1)
                return tb_fail(ctx, TB_ERR_ARG, "tb_correct_download: tracks[t][1] must be NULL unless split_strands");
    if (n > 0)
        for (int t = 0; t < 4; t++)
            for (int s = 0; s < 2; s++)
                if (tracks[t][s]) TB_TRY(tb_reserve(ctx, C.dl[t][s], (size_t)n * esz));
    TB_CUDA(ctx, cudaEventRecord(ctx->ev_copy, ctx->stream));
    TB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy, 0));
    if (n > 0) {
        // all packing kernels first (they run before the caller's next kernels fill the SMs), then the copies back to back
        for (int t = 0; t < 4; t++)
            for (int s = 0; s < 2; s++) {
                if (!tracks[t][s]) continue;
                const double *a = C.tracks.as<double>() + ((i64)t * 2 + (split_strands ? s : 0)) * n;
                const double *b = split_strands ? nullptr : C.tracks.as<double>() + ((i64)t * 2 + 1) * n;
                i64 want = tb_div_up(n, 256), capg = (i64)ctx->sm_count * 16;
                unsigned grid = (unsigned)(want < capg ? want : capg);
                if (as_f64) {
                    auto kp = tb_pack_track_kernel<double>;
                    TB_LAUNCH_ON(ctx, ctx->copy_stream, kp, grid, 256, 0, a, b, n, C.dl[t][s].as<double>());
                } else {
                    auto kp = tb_pack_track_kernel<float>;
                    TB_LAUNCH_ON(ctx, ctx->copy_stream, kp, grid, 256, 0, a, b, n, C.dl[t][s].as<float>());
                }
            }
        for (int t = 0; t < 4; t++)
            for (int s = 0; s < 2; s++)
                if (tracks[t][s])
                    TB_CUDA(ctx, cudaMemcpyAsync(tracks[t][s], C.dl[t][s].p, (size_t)n * esz, cudaMemcpyDeviceToHost, ctx->copy_stream));
    }
    C.pp_dst = prepost_out;
    if (prepost_out) {      // through a pinned staging buffer: a copy into pageable memory would block the host until the stream drains
        if (!C.pp_pinned) TB_CUDA(ctx, cudaMallocHost((void **)&C.pp_pinned, (size_t)2 * 3 * 5 * 32 * sizeof(double)));
        TB_CUDA(ctx, cudaMemcpyAsync(C.pp_pinned, C.prepost.p, (size_t)2 * 3 * 5 * C.L * sizeof(double), cudaMemcpyDeviceToHost,
                                     ctx->copy_stream));
    }
    return tb_check(ctx, cudaGetLastError(), "tb_correct_download_begin");
}

extern "C" int tb_correct_download_wait(tb_ctx *ctx) {
    if (!ctx) return TB_ERR_ARG;
    TB_TRY(tb_check(ctx, cudaStreamSynchronize(ctx->copy_stream), "tb_correct_download_wait"));
    if (ctx->cor.pp_dst) {
        memcpy(ctx->cor.pp_dst, ctx->cor.pp_pinned, (size_t)2 * 3 * 5 * ctx->cor.L * sizeof(double));
        ctx->cor.pp_dst = nullptr;
    }
    return tb_check(ctx, cudaGetLastError(), "tb_correct_download_wait");
}

extern "C" int tb_correct_download(tb_ctx *ctx, int split_strands, int as_f64, void *tracks[4][2], double *prepost_out) {
    TB_TRY(tb_correct_download_begin(ctx, split_strands, as_f64, tracks, prepost_out));
    return tb_correct_download_wait(ctx);
}

extern "C" int tb_correct_fit_trace(tb_ctx *ctx, int64_t *n_windows_total, double *out) {
    if (!ctx) return TB_ERR_ARG;
    auto &C = ctx->cor;
    if (!C.valid) return tb_fail(ctx, TB_ERR_STATE, "tb_correct_fit_trace: no completed tb_correct_run");
    if (n_windows_total) *n_windows_total = C.n_windows;
    if (out && C.n_windows) TB_TRY(tb_d2h(ctx, out, C.fit.p, (size_t)C.n_windows * 2 * TB_FIT_REC * sizeof(double)));
    return TB_OK;
}
