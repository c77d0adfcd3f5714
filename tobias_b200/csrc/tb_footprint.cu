// tb_footprint.cu -- K4: ScoreBigwig scoring (multi-width flank-vs-centre footprint scan, rolling sum / mean, smoothing).
//
// Reference: calculate_scores, tobias/tools/score_bigwig.py:33-98; tobias_footprint_array, tobias/utils/signals.pyx:184-230;
// fast_rolling_math, signals.pyx:102-177; FOS_score, signals.pyx:237-282.
//
// Footprint scan, gather form of the reference's scatter loop.  With P / N the running sums of max(v,0) / min(v,0) over the
// fp32-truncated input, a footprint starting at s with flank width fw and footprint width pw scores
//     score(s,fw,pw) = ( (P[s]-P[s-fw]) + (P[s+pw+fw]-P[s+pw]) ) / (2 fw)  -  (N[s+pw]-N[s]) / pw ,   0 <= s-fw < L-2*fmax-pmax
// and out[pos] = max(0, max_{fw,pw} max_{s in (pos-pw, pos]} score(s,fw,pw)).  Per tile: a block-wide fp64 scan builds P and N in
// shared memory (tile-local, so differences keep full precision); for each fw the fp32 window sums W_fw[k] = P[k+fw]-P[k] are
// staged once in shared memory and every thread (one footprint start s) folds them into 32 register accumulators, one per pw;
// the footprint term is subtracted; a log-doubling sliding maximum in shared memory spreads each score over its footprint.
// The reference sums sequentially in fp32; prefix differences rounded to fp32 agree with it to a few ulp (tests: rtol 1e-5).
#include "tb_common.cuh"

#include <algorithm>
#include <cfloat>
#include <utility>

struct tb_sc_view {
    const i64 *ext_off;      // n+1 offsets of the (flank-extended) regions in the concatenated signal
    const i64 *out_off;      // n+1 offsets of the trimmed outputs
    const i64 *tile_off;     // n+1 prefix of tiles per region
    i64 n_regions;
    int flank;
};

struct tb_prep {
    int absolute, use_min, use_max;
    double min_limit, max_limit;
};

// nan_to_num + --absolute + --min-limit / --max-limit (score_bigwig.py:55-64), evaluated in double like the reference
__device__ __forceinline__ double tb_prep_value(float v, const tb_prep &p) {
    if (v != v) v = 0.0f;
    else if (v > FLT_MAX) v = FLT_MAX;
    else if (v < -FLT_MAX) v = -FLT_MAX;
    double d = (double)v;
    if (p.absolute) d = fabs(d);
    if (p.use_min && d < p.min_limit) d = p.min_limit;
    if (p.use_max && d > p.max_limit) d = p.max_limit;
    return d;
}

// exclusive block scan of two fp64 sequences held in shared memory: a[0..n), b[0..n) -> in place prefix (a[i] = sum of inputs < i),
// a[n], b[n] receive the totals.  aux needs 2 * blockDim.x doubles.
__device__ void tb_block_scan2(double *a, double *b, int n, double *aux) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int per = (n + nt - 1) / nt;
    const int lo = min(n, tid * per), hi = min(n, lo + per);
    double sa = 0.0, sb = 0.0;
    for (int i = lo; i < hi; i++) {
        sa += a[i];
        sb += b[i];
    }
    aux[tid] = sa;
    aux[nt + tid] = sb;
    __syncthreads();
    if (tid < 32) {
        const int chunk = (nt + 31) / 32;
        const int c0 = min(nt, tid * chunk), c1 = min(nt, c0 + chunk);
        double ta = 0.0, tb = 0.0;
        for (int i = c0; i < c1; i++) {
            ta += aux[i];
            tb += aux[nt + i];
        }
        double ia = ta, ib = tb;                                  // inclusive scan over lanes
        for (int d = 1; d < 32; d <<= 1) {
            double oa = __shfl_up_sync(0xffffffffu, ia, d), ob = __shfl_up_sync(0xffffffffu, ib, d);
            if (tid >= d) {
                ia += oa;
                ib += ob;
            }
        }
        double ra = ia - ta, rb = ib - tb;                        // exclusive offset of this lane's chunk
        for (int i = c0; i < c1; i++) {
            double va = aux[i], vb = aux[nt + i];
            aux[i] = ra;
            aux[nt + i] = rb;
            ra += va;
            rb += vb;
        }
    }
    __syncthreads();
    double ra = aux[tid], rb = aux[nt + tid];
    for (int i = lo; i < hi; i++) {
        double va = a[i], vb = b[i];
        a[i] = ra;
        b[i] = rb;
        ra += va;
        rb += vb;
    }
    if (hi == n && lo < n) {
        a[n] = ra;
        b[n] = rb;
    }
    if (n == 0 && tid == 0) {
        a[0] = 0.0;
        b[0] = 0.0;
    }
    __syncthreads();
}

#define TB_K4_THREADS 512
#define TB_K4_NPW 32

// OUT_MODE 0: trimmed float, 1: trimmed double, 2: untrimmed double work array (smoothing follows)
// The block size S (= blockDim.x, chosen by the host to fit the region lengths) sets the tile: S footprint starts,
// S - (pmax - 1) outputs.  W is double buffered (one barrier per flank width).  Spreading a score over its footprint:
// with g[c] = max_{pw >= pmin + c} score(s, pw) (suffix maximum over the widths, in registers),
//     out[pos] = max( max_{d < pmin} g[0](pos - d),  max_{c >= 1} g[c](pos - (pmin - 1 + c)) )
// i.e. one shared-memory exchange per width instead of a log-doubling sliding maximum per width (all widths of one batch
// of 32; wider ranges fall back to the per-width sliding maximum).
template <int OUT_MODE>
__global__ void __launch_bounds__(TB_K4_THREADS)
tb_footprint_kernel(tb_sc_view V, const float *__restrict__ signal, tb_prep prep, int fmin_, int fmax_, int pmin_, int pmax_,
                    void *__restrict__ out_) {
    TB_DYN_SMEM(smem);
    const int tid = threadIdx.x;
    const int S = blockDim.x;
    const int T = S - (pmax_ - 1);                              // outputs per tile
    const i64 j = tb_upper_bound(V.tile_off, V.n_regions + 1, (i64)blockIdx.x) - 1;
    const i64 tile = (i64)blockIdx.x - V.tile_off[j];
    const i64 Lr = V.ext_off[j + 1] - V.ext_off[j];
    const i64 p0 = tile * T;                                    // first output position of the tile (region coordinates)
    const i64 smin = p0 - (pmax_ - 1);                          // footprint start handled by thread 0
    const i64 in0 = smin - fmax_;                               // first loaded position
    const int n_in = S + 2 * fmax_ + pmax_;
    const i64 limit = Lr - 2 * (i64)fmax_ - pmax_;              // i = s - fw must be < limit
    double *P = (double *)smem;                                 // n_in + 1
    double *N = P + (n_in + 1);                                 // n_in + 1
    double *aux = N + (n_in + 1);                               // 2 * S
    float *W0 = (float *)(aux + 2 * S);                         // n_in + 32 (the width loop reads a full batch of 32), two buffers
    float *W1 = W0 + n_in + TB_K4_NPW;
    float *A = W1 + n_in + TB_K4_NPW;                           // S
    float *B = A + S;                                           // S
    const float *sg = signal + V.ext_off[j];

    for (int i = tid; i < n_in; i += S) {
        i64 q = in0 + i;
        float v = 0.0f;
        if (q >= 0 && q < Lr) v = (float)tb_prep_value(sg[q], prep);      // `val = arr[i+j]` is a C float (signals.pyx:191)
        P[i] = v > 0.0f ? (double)v : 0.0;
        N[i] = v < 0.0f ? (double)v : 0.0;
    }
    __syncthreads();
    tb_block_scan2(P, N, n_in, aux);

    const i64 s = smin + tid;                                   // this thread's footprint start
    const int si = tid + fmax_;                                 // its index in the loaded range
    float best = 0.0f;                                          // out[pos] for pos = s (footprint_scores starts at 0)
    const bool one_batch = pmax_ - pmin_ + 1 <= TB_K4_NPW;
    int wbuf = 0;

    for (int pw0 = pmin_; pw0 <= pmax_; pw0 += TB_K4_NPW) {
        const int npw = min(TB_K4_NPW, pmax_ - pw0 + 1);
        float acc[TB_K4_NPW];
#pragma unroll
        for (int c = 0; c < TB_K4_NPW; c++) acc[c] = -INFINITY;
        for (int fw = fmin_; fw <= fmax_; fw++) {
            float *W = wbuf ? W1 : W0;
            wbuf ^= 1;
            // flank sums already divided by 2 * fw (once per element instead of once per element and footprint width)
            const float inv = 1.0f / (2.0f * (float)fw);
            for (int i = tid; i + fw <= n_in; i += S) W[i] = (float)(P[i + fw] - P[i]) * inv;
            __syncthreads();                                    // the other buffer is rewritten only after the next barrier
            const i64 i0 = s - fw;
            if (i0 >= 0 && i0 < limit) {
                const float l = W[si - fw];
                const float *wr = W + si + pw0;
#pragma unroll
                for (int c = 0; c < TB_K4_NPW; c++) acc[c] = fmaxf(acc[c], l + wr[c]);      // c >= npw: padding, dropped below
            }
        }
        // acc[c] -> score of (s, pw0 + c)
#pragma unroll
        for (int c = 0; c < TB_K4_NPW; c++) {
            float sc = -INFINITY;
            if (c < npw && acc[c] > -INFINITY) {
                const int pw = pw0 + c;
                float mid = (float)(N[si + pw] - N[si]);
                sc = acc[c] - mid * (1.0f / (float)pw);
            }
            acc[c] = sc;
        }
        if (one_batch) {
            // suffix maxima over the widths
#pragma unroll
            for (int c = TB_K4_NPW - 2; c >= 0; c--) acc[c] = fmaxf(acc[c], acc[c + 1]);
            __syncthreads();
            A[tid] = acc[0];
            __syncthreads();
            for (int d = 0; d < pmin_ && d <= tid; d++) best = fmaxf(best, A[tid - d]);
            float *cur = B;
#pragma unroll
            for (int c = 1; c < TB_K4_NPW; c++) {
                if (c < npw) {                                  // block-uniform
                    cur[tid] = acc[c];
                    __syncthreads();
                    const int d = pmin_ - 1 + c;
                    if (tid >= d) best = fmaxf(best, cur[tid - d]);
                    cur = cur == A ? B : A;
                }
            }
            __syncthreads();
        } else {
#pragma unroll
            for (int c = 0; c < TB_K4_NPW; c++) {
                if (c < npw) {
                    const int pw = pw0 + c;
                    // sliding maximum over s in (pos - pw, pos]: log-doubling
                    __syncthreads();
                    A[tid] = acc[c];
                    __syncthreads();
                    float *src = A, *dst = B;
                    int span = 1;
                    while (span * 2 <= pw) {
                        float v = src[tid];
                        if (tid >= span) v = fmaxf(v, src[tid - span]);
                        dst[tid] = v;
                        __syncthreads();
                        float *t = src; src = dst; dst = t;
                        span *= 2;
                    }
                    float v = src[tid];
                    int back = pw - span;
                    if (back > 0 && tid >= back) v = fmaxf(v, src[tid - back]);
                    best = fmaxf(best, v);
                }
            }
            __syncthreads();
        }
    }
    // thread tid holds out[pos = s]; valid for pos in [p0, p0 + T)
    if (tid >= pmax_ - 1 && s < Lr) {
        if (OUT_MODE == 2) {
            ((double *)out_)[V.ext_off[j] + s] = (double)best;
        } else if (s >= V.flank && s < Lr - V.flank) {
            i64 o = V.out_off[j] + (s - V.flank);
            if (OUT_MODE == 0) ((float *)out_)[o] = best;
            else ((double *)out_)[o] = (double)best;
        }
    }
}

// ---- register-tiled form of the footprint scan: 4 footprint starts x NW footprint widths per thread ----
// For a fixed flank width the score table acc(s, e) = max_fw ( L_fw[s] + R_fw[e] ) (e = s + pw: end of the footprint) is a max-plus product of
// the left-flank column and the right-flank row, so a thread keeps a register tile of accumulators like a GEMM micro-kernel: starts
// 4t .. 4t + 3 against NW consecutive widths need 4 left values and NW + 3 right values of a flank-width row -- 7 16-byte shared-memory loads
// serve 64 (add, max) pairs where the first form issues one 4-byte load per pair (it is bound by the shared-memory pipe).  Two rows are folded per
// step, acc = max(acc, l0 + r0, l1 + r1): the three-input maximum (FMNMX3) costs what the two-input one does, so the ALU pipe, which binds
// the arithmetic (measured: FMNMX and FMNMX3 2 cycles per warp, FADD 1), does half the work.  All flank-width rows of a tile are staged once
// (row r + 1 from row r by one fp32 add per element: W_{fw+1}[i] = W_fw[i] + max(v[i + fw], 0), sums of non-negative terms, no cancellation),
// so the row loop runs without barriers; NW = 16 runs two passes over the widths (64 accumulators, two blocks per SM).
// Left values sit at 4t + (fmax - fw): their misalignment against 16 bytes is the same for every thread and known per row, so the row step is
// instantiated for the four cases and picks its registers statically.  A start whose left flank would begin before the region exists in no
// row (signals.pyx:199): those row entries are staged as -inf (they are never read as right values by a valid start); starts beyond the last
// valid one are masked per row in the few threads that have any.
// Centre term: mid(u, c) = (N[s_u + pw] - N[s_3]) + (N[s_3] - N[s_u]), two same-signed fp32 numbers (sums of non-positive values), so rounding
// the fp64 prefix differences once per window position instead of once per (start, width) loses nothing.
// Spreading a score over its footprint: suffix maxima over the widths as in the first form, but the anti-diagonals of the (start, width) table are
// folded inside the thread first, so one exchange through shared memory per pass is enough (see the kernel).
#define TB_K4T_MAXT 256
#define TB_K4T_EPT 8             // row elements a thread stages at most

// 16-byte shared-memory load that the compiler may neither split nor narrow to the components in use (a thread's stride is 16 bytes: whole
// vectors are conflict free, scalar loads of single components are 4-way conflicts)
__device__ __forceinline__ float4 tb_lds128(const float *p) {
#ifdef TB_EMU
    return *(const float4 *)p;
#else
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "r"((unsigned)__cvta_generic_to_shared(p)));
    return r;
#endif
}

template <int NW, int RHO, bool TWO>
__device__ __forceinline__ void tb_k4t_step(const float *__restrict__ Wr, int wstride, int lb, int rb, bool edge, int e0, int r,
                                            float (&acc)[4][NW]) {
    // left values of row r: 4 floats starting RHO floats behind the aligned index lb; of row r + 1: one float earlier
    float l0[4], l1[4];
    {
        const float4 a = tb_lds128(Wr + lb);
        float t[8] = {a.x, a.y, a.z, a.w, 0.0f, 0.0f, 0.0f, 0.0f};
        if (RHO != 0) {
            const float4 b = tb_lds128(Wr + lb + 4);
            t[4] = b.x, t[5] = b.y, t[6] = b.z, t[7] = b.w;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) l0[u] = t[RHO + u];
    }
    if (TWO) {
        constexpr int R1 = (RHO + 3) & 3;
        const float *W1 = Wr + wstride;
        const int lb1 = RHO == 0 ? lb - 4 : lb;
        const float4 a = tb_lds128(W1 + lb1);
        float t[8] = {a.x, a.y, a.z, a.w, 0.0f, 0.0f, 0.0f, 0.0f};
        if (R1 != 0) {
            const float4 b = tb_lds128(W1 + lb1 + 4);
            t[4] = b.x, t[5] = b.y, t[6] = b.z, t[7] = b.w;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) l1[u] = t[R1 + u];
    }
    if (edge) {                                                 // starts at or beyond `limit` for this flank width do not exist
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (r <= e0 + u) l0[u] = -INFINITY;
            if (TWO && r + 1 <= e0 + u) l1[u] = -INFINITY;
        }
    }
    float v0[NW + 4], v1[NW + 4];
#pragma unroll
    for (int m = 0; m < NW / 4 + 1; m++) {
        const float4 q = tb_lds128(Wr + rb + 4 * m);
        v0[4 * m] = q.x, v0[4 * m + 1] = q.y, v0[4 * m + 2] = q.z, v0[4 * m + 3] = q.w;
        if (TWO) {
            const float4 p = tb_lds128(Wr + wstride + rb + 4 * m);
            v1[4 * m] = p.x, v1[4 * m + 1] = p.y, v1[4 * m + 2] = p.z, v1[4 * m + 3] = p.w;
        }
    }
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
        for (int c = 0; c < NW; c++) {
            if (TWO) acc[u][c] = fmaxf(fmaxf(acc[u][c], l0[u] + v0[u + c]), l1[u] + v1[u + c]);
            else acc[u][c] = fmaxf(acc[u][c], l0[u] + v0[u + c]);
        }
}

// All rows of one pass.  X0 = X & 3 is the misalignment of row 0's left values; it drops by one per row, so over four rows the steps repeat.
template <int NW, int X0>
__device__ __forceinline__ void tb_k4t_rows(const float *__restrict__ W, int wstride, int t4, int X, int rb, bool edge, int e0, int nfw,
                                            float (&acc)[4][NW]) {
    int r = 0;
    const float *Wr = W;
    int lb = t4 + (X & ~3);                                     // aligned index at or below the left values of row r
    for (; r + 3 < nfw; r += 4) {
        tb_k4t_step<NW, X0, true>(Wr, wstride, lb, rb, edge, e0, r, acc);
        // rows r + 2, r + 3: misalignment X0 - 2; the aligned index steps down when that wraps
        tb_k4t_step<NW, (X0 + 2) & 3, true>(Wr + 2 * wstride, wstride, X0 < 2 ? lb - 4 : lb, rb, edge, e0, r + 2, acc);
        Wr += 4 * wstride;
        lb -= 4;
    }
    if (r + 1 < nfw) {
        tb_k4t_step<NW, X0, true>(Wr, wstride, lb, rb, edge, e0, r, acc);
        r += 2;
        if (r < nfw) tb_k4t_step<NW, (X0 + 2) & 3, false>(Wr + 2 * wstride, wstride, X0 < 2 ? lb - 4 : lb, rb, edge, e0, r, acc);
    } else if (r < nfw) {
        tb_k4t_step<NW, X0, false>(Wr, wstride, lb, rb, edge, e0, r, acc);
    }
}

// Stages the rows: W_r[i] = mean of max(v, 0) over [i, i + fw), halved: sum / (2 fw); row r + 1 from row r by one add per element.  Thread t owns
// the elements t, t + nthr, ... (EPT of them, the last one only where it is inside the row); entries whose window starts before the region are
// -inf from the start and stay so.
template <int EPT>
__device__ __forceinline__ void tb_k4t_stage(const double *__restrict__ P, const float *__restrict__ vpos, float *__restrict__ W, int wstride,
                                             int shift, int n_in, i64 in0, int fmin_, int nfw, const float *__restrict__ invf) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    const bool last_on = tid + (EPT - 1) * nthr < wstride - shift;
    float U[EPT];
#pragma unroll
    for (int m = 0; m < EPT; m++) {
        const int i = tid + m * nthr;
        U[m] = in0 + i < 0 ? -INFINITY : (i + fmin_ <= n_in ? (float)(P[i + fmin_] - P[i]) : 0.0f);
    }
    float *wp = W + shift + tid;
    const float *vp = vpos + fmin_ + tid;
#pragma unroll 3
    for (int r = 0; r < nfw; r++) {
        const float inv = r < 64 ? invf[r] : 1.0f / (2.0f * (float)(fmin_ + r));
#pragma unroll
        for (int m = 0; m < EPT - 1; m++) {
            wp[m * nthr] = U[m] * inv;
            U[m] += vp[m * nthr];
        }
        if (last_on) {
            wp[(EPT - 1) * nthr] = U[EPT - 1] * inv;
            U[EPT - 1] += vp[(EPT - 1) * nthr];
        }
        wp += wstride;
        vp += 1;
    }
}

struct tb_k4t_tile {
    i64 ext0, out0;              // offsets of the region in the concatenated signal / trimmed output
    i64 len;                     // length of the (flank-extended) region
    i64 tile;                    // tile index within the region
};

// tile table from the per-region prefix of tile counts (built on the device: a whole-genome scan has millions of tiles)
__global__ void tb_k4t_tiles_kernel(const i64 *__restrict__ ext_off, const i64 *__restrict__ out_off, const i64 *__restrict__ tile_off,
                                    i64 n_regions, i64 n_tiles, tb_k4t_tile *__restrict__ tiles) {
    const i64 b = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_tiles) return;
    const i64 j = tb_upper_bound(tile_off, n_regions + 1, b) - 1;
    tb_k4t_tile t;
    t.ext0 = ext_off[j];
    t.out0 = out_off[j];
    t.len = ext_off[j + 1] - ext_off[j];
    t.tile = b - tile_off[j];
    tiles[b] = t;
}

#define TB_K4T_LD 8              // input elements a thread loads at most
// padding (floats of -inf) in front of the exchange rows: the furthest thread / start a position reads from
__host__ __device__ __forceinline__ int tb_k4t_epad(int pmax) { return (pmax / 4 + 16) & ~3; }
__host__ __device__ __forceinline__ int tb_k4t_gpad(int pmax) { return (pmax + 8) & ~3; }

// OUT_MODE as above.  Shared memory: N[n_in + 2] doubles | P[n_in + 2] doubles | vpos[wstride + fmax + 8] floats -- the three later hold the
// exchange rows E[NW + 3][nthr + epad] and G0[gpad + S] -- | vneg[wstride] floats (shifted like the rows) | W[nfw][wstride]
template <int OUT_MODE, int NW>
__global__ void __launch_bounds__(TB_K4T_MAXT, NW == 16 ? 2 : 1)
tb_footprint_tile_kernel(const tb_k4t_tile *__restrict__ tiles, int flank, const float *__restrict__ signal, tb_prep prep, int fmin_, int fmax_,
                         int pmin_, int pmax_, int wstride, int pg_bytes, void *__restrict__ out_) {
    TB_DYN_SMEM(smem);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int S = 4 * nthr;                                     // footprint starts per tile
    const int T = S - (pmax_ - 1);                              // outputs per tile
    const int npw = pmax_ - pmin_ + 1, nfw = fmax_ - fmin_ + 1;
    const tb_k4t_tile td = tiles[blockIdx.x];
    const i64 Lr = td.len;
    const i64 p0 = td.tile * T;                                 // first output position of the tile (region coordinates)
    const i64 smin = p0 - (pmax_ - 1);                          // first footprint start of the tile
    const i64 in0 = smin - fmax_;                               // first loaded position
    const int n_in = S + 2 * fmax_ + pmax_;
    const i64 limit = Lr - 2 * (i64)fmax_ - pmax_;              // a start s exists for flank width fw iff 0 <= s - fw < limit
    double *N = (double *)smem;                                 // n_in + 2
    double *P = N + (n_in + 2);                                 // n_in + 2
    float *vpos = (float *)(P + (n_in + 2));                    // max(v, 0) as float, zero padded up to wstride + fmax + 8
    const int epad = tb_k4t_epad(pmax_), gpad = tb_k4t_gpad(pmax_), estride = nthr + epad;
    float *E = (float *)smem;                                   // [NW + 3][estride] once the rows are staged and the prefix sums read
    float *G0 = E + (size_t)(NW + 3) * estride;                 // [gpad + S]
    float *vneg = (float *)(smem + pg_bytes);                   // min(v, 0) as float, shifted like the rows, zero padded up to wstride
    float *W = vneg + wstride;                                  // [nfw][wstride]
    __shared__ double wtot[64];
    __shared__ float ipw[32], invf[64];                         // 1 / pw, 1 / (2 fw)
    const float *sg = signal + td.ext0;
    const int shift = (4 - ((fmax_ + pmin_) & 3)) & 3;          // rows are shifted so that every thread's first right value is 16-byte aligned

    if (tid < 32) ipw[tid] = 1.0f / (float)(pmin_ + tid);
    if (tid < 64) invf[tid] = 1.0f / (2.0f * (float)(fmin_ + tid));
    {
        // all loads of the thread first (independent), then the conversions
        const int n_ld = wstride + fmax_ + 8;
        float v[TB_K4T_LD];
#pragma unroll
        for (int m = 0; m < TB_K4T_LD; m++) {
            const int i = tid + m * nthr;
            const i64 q = in0 + i;
            v[m] = (i < n_in && q >= 0 && q < Lr) ? sg[q] : 0.0f;
        }
#pragma unroll
        for (int m = 0; m < TB_K4T_LD; m++) {
            const int i = tid + m * nthr;
            if (i < n_ld) {
                const float x = (float)tb_prep_value(v[m], prep);                     // `val = arr[i+j]` is a C float (signals.pyx:191)
                if (i < n_in) {
                    P[i] = x > 0.0f ? (double)x : 0.0;
                    N[i] = x < 0.0f ? (double)x : 0.0;
                }
                vpos[i] = x > 0.0f ? x : 0.0f;
                if (i < wstride - shift) vneg[shift + i] = x < 0.0f ? x : 0.0f;
            }
        }
    }
    if (tid < shift) vneg[tid] = 0.0f;
    __syncthreads();
    tb_block_scan2w(P, N, n_in, wtot);

    switch ((wstride - shift + nthr - 1) / nthr) {
    case 1: case 2: case 3: case 4: case 5: tb_k4t_stage<5>(P, vpos, W, wstride, shift, n_in, in0, fmin_, nfw, invf); break;
    case 6: tb_k4t_stage<6>(P, vpos, W, wstride, shift, n_in, in0, fmin_, nfw, invf); break;
    case 7: tb_k4t_stage<7>(P, vpos, W, wstride, shift, n_in, in0, fmin_, nfw, invf); break;
    default: tb_k4t_stage<TB_K4T_EPT>(P, vpos, W, wstride, shift, n_in, in0, fmin_, nfw, invf); break;
    }
    for (int r = 0; r < nfw; r++)
        if (tid < shift) W[(size_t)r * wstride + tid] = 0.0f;

    const int si = 4 * tid + fmax_;                             // index of start s_0 in the loaded range
    const i64 s0 = smin + 4 * tid;
    // right edge: start s_u exists in row r iff s_u - fmin - r < limit  <=>  r > e0 + u
    const i64 ed = s0 - fmin_ - limit;
    const int e0 = (int)(ed < -8 ? -8 : (ed > 100000 ? 100000 : ed));
    const bool edge = e0 + 3 >= 0;
    const int X = fmax_ + shift - fmin_;                        // left values of row r start at float 4 * tid + X - r of the row
    const int k0 = 4 * tid;                                     // index of s_0 among the tile's starts
    // centre term: N[s_u + pw] - N[s_u] = (N[s_u + pw] - N[s_3]) + (N[s_3] - N[s_u]); the first bracket starts from one fp64 difference per
    // pass and grows by fp32 adds of min(v, 0) (same-signed terms)
    float lf[4], mf0[32 / NW];
    {
        const double n3 = N[si + 3];
#pragma unroll
        for (int u = 0; u < 4; u++) lf[u] = (float)(N[si + u] - n3);       // negated: subtracted below
#pragma unroll
        for (int ps = 0; ps < 32 / NW; ps++) mf0[ps] = (float)(N[min(si + pmin_ + ps * NW, n_in)] - n3);
    }
    float best[4] = {0.0f, 0.0f, 0.0f, 0.0f};                   // out[pos] for pos = s_u (footprint_scores starts at 0, signals.pyx:193)
    float carry[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    __syncthreads();

    bool first_exchange = true;
#pragma unroll
    for (int pass = 32 / NW - 1; pass >= 0; pass--) {
        const int c0 = pass * NW;                               // first width index of the pass
        if (c0 < npw) {                                         // block-uniform
        float acc[4][NW];
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int c = 0; c < NW; c++) acc[u][c] = -INFINITY;
        const int rb = 4 * tid + fmax_ + pmin_ + shift + c0;
        switch (X & 3) {
        case 0: tb_k4t_rows<NW, 0>(W, wstride, 4 * tid, X, rb, edge, e0, nfw, acc); break;
        case 1: tb_k4t_rows<NW, 1>(W, wstride, 4 * tid, X, rb, edge, e0, nfw, acc); break;
        case 2: tb_k4t_rows<NW, 2>(W, wstride, 4 * tid, X, rb, edge, e0, nfw, acc); break;
        default: tb_k4t_rows<NW, 3>(W, wstride, 4 * tid, X, rb, edge, e0, nfw, acc); break;
        }
        // centre term: mf[k] = N[s_0 + pmin + c0 + k] - N[s_3]
        float mf[NW + 4];
        mf[0] = pass == 0 ? mf0[0] : mf0[32 / NW - 1];
#pragma unroll
        for (int m = 0; m < NW / 4 + 1; m++) {
            const float4 q = tb_lds128(vneg + rb + 4 * m);
            if (4 * m + 1 < NW + 4) mf[4 * m + 1] = mf[4 * m] + q.x;
            if (4 * m + 2 < NW + 4) mf[4 * m + 2] = mf[4 * m + 1] + q.y;
            if (4 * m + 3 < NW + 4) mf[4 * m + 3] = mf[4 * m + 2] + q.z;
            if (4 * m + 4 < NW + 4) mf[4 * m + 4] = mf[4 * m + 3] + q.w;
        }
        // Scores, from the widest width down; g[u][c] = max_{pw >= pmin + c} score(s_u, pw) is carried across the widths and passes.  Spreading
        // a score over its footprint:
        //     out[pos] = max( max_{d < pmin} g[0](pos - d),  max_{c >= 1} g[c](pos - (pmin - 1 + c)) ).
        // The second term is a maximum along anti-diagonals of the (start, width) table; the part of every anti-diagonal inside a thread's
        // four starts is folded in registers first, e[t] = max_{u + c = t} g[c](s_u), which belongs to the position s_0 + pmin - 1 + t:
        // NW + 3 values per thread and pass go through shared memory instead of 4 NW, in one exchange (two barriers) per pass.
        float e[NW + 3];
#pragma unroll
        for (int t = 0; t < NW + 3; t++) e[t] = -INFINITY;
#pragma unroll
        for (int cc = NW - 1; cc >= 0; cc--) {
            if (c0 + cc < npw) {                                // block-uniform
                const float ip = ipw[(c0 + cc) & 31];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const float sc = fmaf(lf[u] - mf[u + cc], ip, acc[u][cc]);
                    if (c0 + cc >= 1) e[u + cc] = fmaxf(fmaxf(e[u + cc], carry[u]), sc);
                    carry[u] = fmaxf(carry[u], sc);
                }
            }
        }
        __syncthreads();                                        // the prefix sums / the previous pass's rows no longer read
        if (first_exchange) {
            for (int i = tid; i < (NW + 3) * epad; i += nthr) E[(size_t)(i / epad) * estride + i % epad] = -INFINITY;
            for (int i = tid; i < gpad; i += nthr) G0[i] = -INFINITY;
            first_exchange = false;
        }
#pragma unroll
        for (int t = 0; t < NW + 3; t++) E[(size_t)t * estride + epad + tid] = e[t];
        if (pass == 0) *(float4 *)(G0 + gpad + k0) = make_float4(carry[0], carry[1], carry[2], carry[3]);
        __syncthreads();
        {
            // position s_0 + u takes e[t] of the thread q before this one for t = 4 q + u - (pmin - 1)
            const int base = pmin_ - 1 + c0;                    // t = c0 + t', t' the row of E
            const int qlo = max(0, (base - 3) >> 2), qhi = (base + NW + 2) >> 2;
            for (int q = qlo; q <= qhi; q++) {
                const float *er = E + epad + tid - q;
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int tp = 4 * q + u - base;
                    if (tp >= 0 && tp < NW + 3) best[u] = fmaxf(best[u], er[(size_t)tp * estride]);
                }
            }
        }
        if (pass == 0) {
            // width index 0: window maximum over the pmin starts pos - (pmin - 1) .. pos
            for (int jq = 0; 4 * jq - 3 < pmin_; jq++) {
                const float4 g = *(const float4 *)(G0 + gpad + k0 - 4 * jq);     // starts k0 - 4 jq + (0 .. 3)
                const float ge[4] = {g.x, g.y, g.z, g.w};
                if (jq >= 1 && 4 * jq + 3 < pmin_) {
                    const float m4 = fmaxf(fmaxf(g.x, g.y), fmaxf(g.z, g.w));
#pragma unroll
                    for (int u = 0; u < 4; u++) best[u] = fmaxf(best[u], m4);
                } else {
#pragma unroll
                    for (int u = 0; u < 4; u++)
#pragma unroll
                        for (int x = 0; x < 4; x++) {
                            const int dist = u + 4 * jq - x;
                            if (dist >= 0 && dist < pmin_) best[u] = fmaxf(best[u], ge[x]);
                        }
                }
            }
        }
        }
    }
    // this thread holds out[pos] for pos = s_0 .. s_0 + 3; valid from the tile's first output position on
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const i64 s = s0 + u;
        if (k0 + u >= pmax_ - 1 && s < Lr) {
            if (OUT_MODE == 2) {
                ((double *)out_)[td.ext0 + s] = (double)best[u];
            } else if (s >= flank && s < Lr - flank) {
                const i64 o = td.out0 + (s - flank);
                if (OUT_MODE == 0) ((float *)out_)[o] = best[u];
                else ((double *)out_)[o] = (double)best[u];
            }
        }
    }
}

// FOS_score (signals.pyx:237-282): exclusive width ranges, plain sums, running minimum from 10000.  One thread per output
// position; tile-local fp64 prefix sums.  Not reachable from the reference CLI (parsers.py:84); kept for API parity.
template <int OUT_MODE>
__global__ void __launch_bounds__(TB_K4_THREADS)
tb_fos_kernel(tb_sc_view V, const float *__restrict__ signal, tb_prep prep, int fmin_, int fmax_, int pmin_, int pmax_,
              void *__restrict__ out_) {
    TB_DYN_SMEM(smem);
    const int tid = threadIdx.x;
    const int S = TB_K4_THREADS;
    const int T = S;
    const i64 j = tb_upper_bound(V.tile_off, V.n_regions + 1, (i64)blockIdx.x) - 1;
    const i64 tile = (i64)blockIdx.x - V.tile_off[j];
    const i64 Lr = V.ext_off[j + 1] - V.ext_off[j];
    const i64 p0 = tile * T;
    const i64 in0 = p0 - (pmax_ - 1) - fmax_;
    const int n_in = S + 2 * (fmax_ + pmax_);
    const i64 limit = Lr - 2 * (i64)fmax_ - pmax_;
    double *P = (double *)smem;
    double *N = P + (n_in + 1);
    double *aux = N + (n_in + 1);
    const float *sg = signal + V.ext_off[j];
    for (int i = tid; i < n_in; i += S) {
        i64 q = in0 + i;
        double v = 0.0;
        if (q >= 0 && q < Lr) v = tb_prep_value(sg[q], prep);
        P[i] = v;
        N[i] = 0.0;
    }
    __syncthreads();
    tb_block_scan2(P, N, n_in, aux);
    const i64 pos = p0 + tid;
    float best = 10000.0f;
    if (pos < Lr) {
        for (int fw = fmin_; fw < fmax_; fw++)
            for (int pw = pmin_; pw < pmax_; pw++)
                for (i64 s = pos - pw + 1; s <= pos; s++) {
                    i64 i0 = s - fw;
                    if (i0 < 0 || i0 >= limit) continue;
                    int b = (int)(s - in0);
                    float left = (float)(P[b] - P[b - fw]), mid = (float)(P[b + pw] - P[b]);
                    float right = (float)(P[b + pw + fw] - P[b + pw]);
                    float Lm = (float)(left / (fw * 1.0)), Cm = (float)(mid / (pw * 1.0)), Rm = (float)(right / (fw * 1.0));
                    float sc = 10000.0f;
                    if (Cm < Rm && Cm < Lm && Lm > 0 && Rm > 0) sc = (float)((Cm + 1.0) / Lm + (Cm + 1.0) / Rm);
                    if (sc < best) best = sc;
                }
        if (OUT_MODE == 2) ((double *)out_)[V.ext_off[j] + pos] = (double)best;
        else if (pos >= V.flank && pos < Lr - V.flank) {
            i64 o = V.out_off[j] + (pos - V.flank);
            if (OUT_MODE == 0) ((float *)out_)[o] = best;
            else ((double *)out_)[o] = (double)best;
        }
    }
}

// fast_rolling_math "sum" / "mean" (signals.pyx:149-164): the accumulator is a C float, each += a double add rounded to float.
// IN_MODE 0: float signal + preprocessing, 1: double array.  is_mean: divide by w (in double).  w == 0: copy (--score none).
#define TB_ROLL_TILE 1024
template <int IN_MODE, int OUT_MODE>
__global__ void tb_rolling_kernel(tb_sc_view V, const void *__restrict__ in_, tb_prep prep, int w, int is_mean,
                                  void *__restrict__ out_) {
    TB_DYN_SMEM(smem);
    double *buf = (double *)smem;                                // TB_ROLL_TILE + w
    const i64 j = tb_upper_bound(V.tile_off, V.n_regions + 1, (i64)blockIdx.x) - 1;
    const i64 tile = (i64)blockIdx.x - V.tile_off[j];
    const i64 Lr = V.ext_off[j + 1] - V.ext_off[j];
    const i64 p0 = tile * TB_ROLL_TILE;
    const int lf = w / 2;
    const i64 in0 = p0 - lf;
    const int n_in = TB_ROLL_TILE + (w > 0 ? w : 0);
    for (int i = threadIdx.x; i < n_in; i += blockDim.x) {
        i64 q = in0 + i;
        double v = 0.0;
        if (q >= 0 && q < Lr) {
            if (IN_MODE == 0) v = tb_prep_value(((const float *)in_)[V.ext_off[j] + q], prep);
            else v = ((const double *)in_)[V.ext_off[j] + q];
        }
        buf[i] = v;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < TB_ROLL_TILE; t += blockDim.x) {
        i64 pos = p0 + t;
        if (pos >= Lr) break;
        double r;
        if (w <= 0) r = buf[t];
        else if (is_mean >= 2) {
            // "max" / "min" / "prod" (signals.pyx:118-146, 166-175): defined at every position; the window is clipped to 0 < index < L
            // (index 0 is never looked at, as in the reference); maxval / minval are C floats (:116), prod is a double
            const double centre = buf[t + lf];
            float ext = (float)centre;
            double prod = 1.0;
            for (int u = 0; u < w; u++) {
                const i64 idx = pos - lf + u;
                if (idx > 0 && idx < Lr) {
                    const double v = buf[t + u];
                    if (is_mean == 2) {
                        if (v > (double)ext) ext = (float)v;
                    } else if (is_mean == 3) {
                        if (v < (double)ext) ext = (float)v;
                    } else prod *= v;
                }
            }
            r = is_mean == 4 ? prod : (double)ext;
        }
        else if (pos >= lf && pos - lf + w <= Lr) {
            float acc = 0.0f;
            for (int u = 0; u < w; u++) acc = (float)((double)acc + buf[t + u]);
            r = is_mean ? (double)acc / (w * 1.0) : (double)acc;
        } else r = (double)NAN;
        if (OUT_MODE == 2) ((double *)out_)[V.ext_off[j] + pos] = r;
        else if (pos >= V.flank && pos < Lr - V.flank) {
            i64 o = V.out_off[j] + (pos - V.flank);
            if (OUT_MODE == 0) ((float *)out_)[o] = (float)r;
            else ((double *)out_)[o] = r;
        }
    }
}

// The corrected track of the last tb_correct_run, read the way ScoreBigwig reads the corrected bigWig (score_bigwig.py:52-53): every
// position of the scoring regions (already extended by `flank`) takes the float32 value of the output region that covers it -- its own
// or an abutting neighbour (50-kb split pieces) -- and 0 where the track holds nothing (nan_to_num of the bigWig's NaN).
// sc_lstart: linear start of every extended scoring region; out_lstart: linear start of every output region (ascending).
// A warp takes runs of 1024 consecutive positions: two lanes find the scoring regions and the output regions the run can touch (binary
// searches over the whole tables, once per run); every position then searches those few entries only.
#define TB_C2S_RUN 1024
__global__ void tb_corrected_to_signal_kernel(const double *__restrict__ cf, const double *__restrict__ cr, const i64 *__restrict__ out_lstart,
                                              const i64 *__restrict__ out_off, i64 n_out, const i64 *__restrict__ sc_lstart,
                                              const i64 *__restrict__ sig_off, i64 n_regions, i64 total, float *__restrict__ sig) {
    const int lane = threadIdx.x & 31;
    const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    for (i64 x0 = warp * TB_C2S_RUN; x0 < total; x0 += nwarps * TB_C2S_RUN) {
        const i64 x1 = (x0 + TB_C2S_RUN < total ? x0 + TB_C2S_RUN : total) - 1;
        i64 rr = 0;
        if (lane < 2) rr = tb_upper_bound(sig_off, n_regions + 1, lane == 0 ? x0 : x1) - 1;
        const i64 r0 = __shfl_sync(0xffffffffu, rr, 0), r1 = __shfl_sync(0xffffffffu, rr, 1);
        // output regions between the one covering the first scoring region's start and the one covering the run's last position
        i64 jj = 0;
        if (lane == 0) jj = tb_upper_bound(out_lstart, n_out, sc_lstart[r0]) - 1;
        if (lane == 1) jj = tb_upper_bound(out_lstart, n_out, sc_lstart[r1] + (x1 - sig_off[r1])) - 1;
        i64 j0 = __shfl_sync(0xffffffffu, jj, 0);
        const i64 j1 = __shfl_sync(0xffffffffu, jj, 1);
        if (j0 < 0) j0 = 0;
        for (i64 x = x0 + lane; x <= x1; x += 32) {
            const i64 r = r0 + tb_upper_bound(sig_off + r0 + 1, r1 - r0, x);
            const i64 lin = sc_lstart[r] + (x - sig_off[r]);
            i64 j = -1;
            if (j1 >= j0) {
                j = j0 + tb_upper_bound(out_lstart + j0, j1 - j0 + 1, lin) - 1;
                // scoring regions out of order: the bracket does not hold, search everything
                if ((j < j0 && j0 > 0) || (j == j1 && j1 + 1 < n_out && out_lstart[j1 + 1] <= lin)) j = tb_upper_bound(out_lstart, n_out, lin) - 1;
            } else if (n_out > 0 && out_lstart[0] <= lin) {
                j = tb_upper_bound(out_lstart, n_out, lin) - 1;
            }
            float v = 0.0f;
            if (j >= 0) {
                const i64 q = lin - out_lstart[j];
                if (q < out_off[j + 1] - out_off[j]) v = (float)(cf[out_off[j] + q] + cr[out_off[j] + q]);
            }
            sig[x] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------- host side
static int tb_sc_layout(tb_ctx *ctx, i64 n_regions, const int64_t *ext_len) {
    auto &S = ctx->sc;
    S.n_regions = n_regions;
    S.ext_off.assign(n_regions + 1, 0);
    for (i64 i = 0; i < n_regions; i++) {
        if (ext_len[i] <= 0) return tb_fail(ctx, TB_ERR_ARG, "region %lld has length %lld", i, (i64)ext_len[i]);
        S.ext_off[i + 1] = S.ext_off[i] + ext_len[i];
    }
    S.ext_total = S.ext_off[n_regions];
    S.have_out = false;
    S.flank = -1;
    return tb_h2d(ctx, S.d_ext_off, S.ext_off.data(), sizeof(i64) * (n_regions + 1));
}

extern "C" int tb_signal_upload(tb_ctx *ctx, const float *signal, int64_t n_regions, const int64_t *ext_len) {
    if (!ctx || (n_regions > 0 && (!signal || !ext_len)) || n_regions < 0) return TB_ERR_ARG;
    ctx->sc.have_signal = false;
    TB_TRY(tb_sc_layout(ctx, n_regions, ext_len));
    TB_TRY(tb_h2d(ctx, ctx->sc.signal, signal, (size_t)ctx->sc.ext_total * sizeof(float)));
    TB_TRY(tb_sync(ctx, "tb_signal_upload"));
    ctx->sc.have_signal = true;
    return TB_OK;
}

static int tb_signal_from_corrected_linear(tb_ctx *ctx, int flank, i64 n, const std::vector<i64> &sc_lstart, const std::vector<int64_t> &ext_len) {
    auto &C = ctx->cor;
    ctx->sc.have_signal = false;
    const tb_tracer tr(ctx, "tb_signal_from_corrected");
    TB_TRY(tb_sc_layout(ctx, n, ext_len.data()));
    auto &S = ctx->sc;
    TB_TRY(tb_reserve(ctx, S.signal, (size_t)S.ext_total * sizeof(float)));
    if (S.ext_total > 0) {
        std::vector<i64> out_lstart(C.n_regions);
        for (i64 i = 0; i < C.n_regions; i++) out_lstart[i] = C.ext_start[i] + C.k_flank + C.window / 2;
        TB_TRY(tb_h2d(ctx, ctx->reg_a, out_lstart.data(), sizeof(i64) * (size_t)C.n_regions));
        TB_TRY(tb_h2d(ctx, ctx->reg_b, sc_lstart.data(), sizeof(i64) * (size_t)n));
        tr.mark("tables uploaded");
        i64 want = tb_div_up(S.ext_total, (i64)TB_C2S_RUN * 8), capg = (i64)ctx->sm_count * 16;
        const double *cf = C.tracks.as<double>() + ((i64)3 * 2 + 0) * C.out_total;
        const double *cr = C.tracks.as<double>() + ((i64)3 * 2 + 1) * C.out_total;
        TB_LAUNCH(ctx, tb_corrected_to_signal_kernel, (unsigned)(want < capg ? want : capg), 256, 0, cf, cr, (const i64 *)ctx->reg_a.as<i64>(),
                  (const i64 *)C.d_out_off.as<i64>(), C.n_regions, (const i64 *)ctx->reg_b.as<i64>(), (const i64 *)S.d_ext_off.as<i64>(), n,
                  S.ext_total, S.signal.as<float>());
    }
    TB_TRY(tb_sync(ctx, "tb_signal_from_corrected"));
    S.have_signal = true;
    return TB_OK;
}

extern "C" int tb_signal_from_corrected(tb_ctx *ctx, int flank) {
    if (!ctx) return TB_ERR_ARG;
    auto &C = ctx->cor;
    if (!C.valid) return tb_fail(ctx, TB_ERR_STATE, "tb_signal_from_corrected: no completed tb_correct_run");
    if (flank < 0) return tb_fail(ctx, TB_ERR_ARG, "tb_signal_from_corrected: negative flank");
    std::vector<int64_t> ext_len(C.n_regions);
    std::vector<i64> sc_lstart(C.n_regions);
    for (i64 i = 0; i < C.n_regions; i++) {
        ext_len[i] = (C.out_off[i + 1] - C.out_off[i]) + 2 * (i64)flank;
        sc_lstart[i] = C.ext_start[i] + C.k_flank + C.window / 2 - flank;
    }
    return tb_signal_from_corrected_linear(ctx, flank, C.n_regions, sc_lstart, ext_len);
}

extern "C" int tb_signal_from_corrected_regions(tb_ctx *ctx, int flank, int64_t n_regions, const int32_t *contig, const int64_t *start,
                                                const int64_t *end) {
    if (!ctx || n_regions < 0 || (n_regions > 0 && (!contig || !start || !end))) return TB_ERR_ARG;
    auto &C = ctx->cor;
    if (!C.valid) return tb_fail(ctx, TB_ERR_STATE, "tb_signal_from_corrected_regions: no completed tb_correct_run");
    if (flank < 0) return tb_fail(ctx, TB_ERR_ARG, "tb_signal_from_corrected_regions: negative flank");
    std::vector<i64> ls, le;
    TB_TRY(tb_linear_regions(ctx, n_regions, contig, start, end, ls, le, false));
    std::vector<int64_t> ext_len(n_regions);
    std::vector<i64> sc_lstart(n_regions);
    for (i64 i = 0; i < n_regions; i++) {
        if (start[i] - flank < 0 || end[i] + flank > ctx->contig_len[contig[i]])
            return tb_fail(ctx, TB_ERR_ARG, "tb_signal_from_corrected_regions: region %lld extended by %d leaves contig %d", i, flank, contig[i]);
        ext_len[i] = (le[i] - ls[i]) + 2 * (i64)flank;
        sc_lstart[i] = ls[i] - flank;
    }
    return tb_signal_from_corrected_linear(ctx, flank, n_regions, sc_lstart, ext_len);
}

template <int IN_MODE>
static int tb_launch_rolling(tb_ctx *ctx, const tb_sc_view &V, i64 n_tiles, const void *in, const tb_prep &prep, int w, int is_mean,
                             int out_mode, void *out) {
    size_t smem = (size_t)(TB_ROLL_TILE + (w > 0 ? w : 0)) * sizeof(double);
    if (smem > (size_t)200 * 1024) return tb_fail(ctx, TB_ERR_ARG, "rolling window %d too large", w);
#define TB_ROLL_CASE(OM)                                                                                        \
    {                                                                                                           \
        auto k = tb_rolling_kernel<IN_MODE, OM>;                                                                \
        TB_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));          \
        TB_LAUNCH(ctx, k, (unsigned)n_tiles, 256, smem, V, in, prep, w, is_mean, out);                          \
    }
    if (out_mode == 0) TB_ROLL_CASE(0) else if (out_mode == 1) TB_ROLL_CASE(1) else TB_ROLL_CASE(2)
#undef TB_ROLL_CASE
    return tb_check(ctx, cudaGetLastError(), "tb_rolling_kernel");
}

static int tb_tiles(tb_ctx *ctx, const std::vector<i64> &ext_off, i64 per_tile, tb_dbuf &d_tile_off, i64 &n_tiles,
                    std::vector<i64> *host_tile_off = nullptr) {
    i64 n = (i64)ext_off.size() - 1;
    std::vector<i64> tile_off(n + 1, 0);
    for (i64 i = 0; i < n; i++) tile_off[i + 1] = tile_off[i] + tb_div_up(ext_off[i + 1] - ext_off[i], per_tile);
    n_tiles = tile_off[n];
    if (n_tiles > 0x7fffffffLL) return tb_fail(ctx, TB_ERR_ARG, "too many tiles (%lld)", n_tiles);
    TB_TRY(tb_h2d(ctx, d_tile_off, tile_off.data(), sizeof(i64) * (n + 1)));
    if (host_tile_off) host_tile_off->swap(tile_off);
    return TB_OK;
}

extern "C" int tb_score_run(tb_ctx *ctx, const tb_score_params *p) {
    if (!ctx || !p) return TB_ERR_ARG;
    auto &S = ctx->sc;
    if (!S.have_signal) return tb_fail(ctx, TB_ERR_STATE, "tb_score_run: no signal (tb_signal_upload / tb_signal_from_corrected)");
    const tb_tracer tr(ctx, "tb_score_run");
    void *const sink = S.sink;                                   // serves this run only, whatever its outcome
    S.sink = nullptr;
    bool sink_served = false;
    S.have_out = false;
    if (p->flank < 0) return tb_fail(ctx, TB_ERR_ARG, "tb_score_run: negative flank");
    const bool fp_like = p->kind == TB_SCORE_FOOTPRINT || p->kind == TB_SCORE_FOS;
    if (fp_like && (p->flank_min < 1 || p->flank_max < p->flank_min || p->fp_min < 1 || p->fp_max < p->fp_min || p->flank_max > 2048 ||
                    p->fp_max > 256))
        return tb_fail(ctx, TB_ERR_ARG, "tb_score_run: need 1 <= flank_min <= flank_max <= 2048 and 1 <= fp_min <= fp_max <= 256");
    if ((p->kind == TB_SCORE_SUM || p->kind == TB_SCORE_MEAN) && p->window < 1)
        return tb_fail(ctx, TB_ERR_ARG, "tb_score_run: window must be >= 1");
    S.out_off.assign(S.n_regions + 1, 0);
    for (i64 i = 0; i < S.n_regions; i++) {
        i64 len = S.ext_off[i + 1] - S.ext_off[i] - 2 * (i64)p->flank;
        if (len <= 0) return tb_fail(ctx, TB_ERR_ARG, "tb_score_run: region %lld shorter than twice the flank", i);
        S.out_off[i + 1] = S.out_off[i] + len;
    }
    S.out_total = S.out_off[S.n_regions];
    S.out_f64 = p->as_f64 ? 1 : 0;
    S.flank = p->flank;
    if (S.n_regions == 0) {
        S.have_out = true;
        return TB_OK;
    }
    TB_TRY(tb_h2d(ctx, ctx->reg_d, S.out_off.data(), sizeof(i64) * (S.n_regions + 1)));
    TB_TRY(tb_reserve(ctx, S.out, (size_t)S.out_total * (p->as_f64 ? 8 : 4)));
    const bool smooth = p->smooth > 1;
    if (smooth) TB_TRY(tb_reserve(ctx, S.work0, (size_t)S.ext_total * sizeof(double)));
    tb_prep prep{p->absolute, p->use_min, p->use_max, p->min_limit, p->max_limit};
    const int first_mode = smooth ? 2 : (p->as_f64 ? 1 : 0);
    void *first_out = smooth ? S.work0.p : S.out.p;
    i64 n_tiles = 0;

    tr.mark("tables, buffers");
    tb_timer_start(ctx);
    if (fp_like) {
        const bool fos = p->kind == TB_SCORE_FOS;
        std::vector<std::pair<i64, i64>> hist;               // (region length, count): most output regions have the same length
        for (i64 i = 0; i < S.n_regions; i++) {
            const i64 v = S.ext_off[i + 1] - S.ext_off[i];
            if (!hist.empty() && hist.back().first == v) hist.back().second++;
            else hist.push_back({v, 1});
        }
        const int npw = p->fp_max - p->fp_min + 1, nfw = p->flank_max - p->flank_min + 1;
        // register-tiled form (4 starts x 16 or 32 widths per thread, all flank-width rows staged): at most 32 footprint widths, and the rows
        // must fit; option k4_form: 0 = first form, 16 / 32 = widths per pass of the tiled form (default 16: two blocks per SM)
        int bs2 = 0, wstride2 = 0, pg2 = 0, nw2 = 0;
        size_t smem2 = 0;
        const int form = (int)tb_opt(ctx, "k4_form", 16);
        if (!fos && npw <= 32 && (form == 16 || form == 32)) {
            double best_cost = -1.0;
            const int forced = (int)tb_opt(ctx, "k4_threads", 0);                                // comparison runs: threads per block
            for (int cand = 64; cand <= TB_K4T_MAXT; cand += 32) {
                if (forced >= 64 && cand != (forced & ~31)) continue;
                const int Sx = 4 * cand;
                const i64 per = Sx - (p->fp_max - 1);
                if (per < 32) continue;
                const int n_in = Sx + 2 * p->flank_max + p->fp_max;
                const int wstride = (Sx + p->flank_max + p->fp_min + 40 + 3) & ~3;         // the last index a thread reads, incl. widths beyond pmax
                if (wstride > TB_K4T_EPT * cand || wstride + p->flank_max + 8 > TB_K4T_LD * cand) continue;
                const size_t xch = (size_t)4 * ((size_t)(form + 3) * (cand + tb_k4t_epad(p->fp_max)) + tb_k4t_gpad(p->fp_max) + Sx);
                const size_t pg = (std::max((size_t)16 * (n_in + 2) + (size_t)4 * (wstride + p->flank_max + 8), xch) + 15) & ~(size_t)15;
                const size_t need = pg + (size_t)4 * wstride + (size_t)nfw * wstride * 4 + 16;
                if (need > (size_t)220 * 1024) continue;
                const int by_smem = (int)(((size_t)227 * 1024) / (need + 1024 + 400));
                const int by_regs = form == 16 ? 65536 / (128 * cand) : 65536 / (255 * cand);
                const int bps = std::max(1, std::min(std::min(by_smem, by_regs), 2048 / cand));
                double cost = 0.0;
                for (const auto &h : hist) cost += (double)tb_div_up(h.first, per) * (double)h.second;
                // work of a tile ~ its slots incl. the halo; few warps per SM hide the staging / exchange phases badly
                cost *= ((double)n_in + 64.0) * (bps * cand >= 384 ? 1.0 : 1.3);
                if (best_cost < 0.0 || cost < best_cost) {
                    best_cost = cost;
                    bs2 = cand;
                    smem2 = need;
                    wstride2 = wstride;
                    pg2 = (int)pg;
                    nw2 = form;
                }
            }
        }
        if (bs2 > 0) {
            const i64 per_tile = 4 * bs2 - (p->fp_max - 1);
            std::vector<i64> tile_off;
            TB_TRY(tb_tiles(ctx, S.ext_off, per_tile, ctx->reg_e, n_tiles, &tile_off));
            TB_TRY(tb_reserve(ctx, ctx->scratch4, sizeof(tb_k4t_tile) * (size_t)std::max<i64>(n_tiles, 1)));
            TB_LAUNCH(ctx, tb_k4t_tiles_kernel, (unsigned)tb_div_up(std::max<i64>(n_tiles, 1), 256), 256, 0, S.d_ext_off.as<i64>(), ctx->reg_d.as<i64>(),
                      ctx->reg_e.as<i64>(), S.n_regions, n_tiles, (tb_k4t_tile *)ctx->scratch4.p);
            // With a sink (tb_score_set_sink; no smoothing pass) the tiles go in chunks that end on region borders, and every chunk's scores
            // travel to the host on the copy stream while the next chunk computes.
            const bool stream_out = sink && !smooth;
            const int n_chunks = stream_out ? (int)std::max<i64>(1, std::min<i64>(tb_opt(ctx, "k4_chunks", 3), std::min<i64>(8, S.n_regions))) : 1;
            const size_t esz = p->as_f64 ? 8 : 4;
#define TB_FP2_CASE(OM, NWX)                                                                                                 \
    {                                                                                                                       \
        auto k = tb_footprint_tile_kernel<OM, NWX>;                                                                         \
        TB_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));                     \
        TB_LAUNCH(ctx, k, (unsigned)cnt, bs2, smem2, (const tb_k4t_tile *)ctx->scratch4.p + t_done, p->flank,              \
                  (const float *)S.signal.as<float>(), prep, p->flank_min, p->flank_max, p->fp_min, p->fp_max, wstride2, pg2, first_out); \
    }
            i64 r_done = 0;
            for (int ch = 0; ch < n_chunks; ch++) {
                const i64 r_end = ch + 1 == n_chunks ? S.n_regions : (S.n_regions * (ch + 1)) / n_chunks;
                if (r_end <= r_done) continue;
                const i64 t_done = tile_off[r_done], cnt = tile_off[r_end] - t_done;
                if (cnt > 0) {
                    if (nw2 == 16) {
                        if (first_mode == 0) TB_FP2_CASE(0, 16) else if (first_mode == 1) TB_FP2_CASE(1, 16) else TB_FP2_CASE(2, 16)
                    } else {
                        if (first_mode == 0) TB_FP2_CASE(0, 32) else if (first_mode == 1) TB_FP2_CASE(1, 32) else TB_FP2_CASE(2, 32)
                    }
                }
                if (stream_out) {
                    const i64 o0 = S.out_off[r_done], o1 = S.out_off[r_end];
                    TB_CUDA(ctx, cudaEventRecord(ctx->ev_chunk[ch & 7], ctx->stream));
                    TB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_chunk[ch & 7], 0));
                    if (o1 > o0)
                        TB_CUDA(ctx, cudaMemcpyAsync((char *)sink + (size_t)o0 * esz, (const char *)S.out.p + (size_t)o0 * esz, (size_t)(o1 - o0) * esz,
                                                     cudaMemcpyDeviceToHost, ctx->copy_stream));
                    sink_served = true;
                }
                r_done = r_end;
            }
#undef TB_FP2_CASE
        } else {
        // footprint scan, first form: block size = tile size, picked to waste the fewest thread slots on these region lengths
        int bs = TB_K4_THREADS;
        if (!fos) {
            double best_cost = -1.0;
            for (int cand = 128; cand <= TB_K4_THREADS; cand += 32) {
                const i64 per = cand - (p->fp_max - 1);
                if (per < 32) continue;
                double cost = 0.0;
                for (const auto &h : hist) cost += (double)tb_div_up(h.first, per) * (double)h.second;
                cost *= (double)(cand + 2 * p->flank_max + p->fp_max);         // slots incl. the halo every tile re-scans
                if (best_cost < 0.0 || cost < best_cost) {
                    best_cost = cost;
                    bs = cand;
                }
            }
        }
        const i64 per_tile = fos ? bs : bs - (p->fp_max - 1);
        TB_TRY(tb_tiles(ctx, S.ext_off, per_tile, ctx->reg_e, n_tiles));
        tb_sc_view V{S.d_ext_off.as<i64>(), ctx->reg_d.as<i64>(), ctx->reg_e.as<i64>(), S.n_regions, p->flank};
        const int n_in = fos ? bs + 2 * (p->flank_max + p->fp_max) : bs + 2 * p->flank_max + p->fp_max;
        size_t smem = (size_t)(2 * (n_in + 1) + 2 * bs) * sizeof(double) + (size_t)(2 * (n_in + TB_K4_NPW) + 2 * bs) * sizeof(float);
        if (smem > (size_t)220 * 1024) return tb_fail(ctx, TB_ERR_ARG, "tb_score_run: flank/footprint widths too large for one tile");
#define TB_FP_CASE(KERNEL, OM)                                                                                              \
    {                                                                                                                       \
        auto k = KERNEL<OM>;                                                                                                \
        TB_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                      \
        TB_LAUNCH(ctx, k, (unsigned)n_tiles, bs, smem, V, (const float *)S.signal.as<float>(), prep, p->flank_min, \
                  p->flank_max, p->fp_min, p->fp_max, first_out);                                                           \
    }
        if (!fos) {
            if (first_mode == 0) TB_FP_CASE(tb_footprint_kernel, 0) else if (first_mode == 1) TB_FP_CASE(tb_footprint_kernel, 1)
            else TB_FP_CASE(tb_footprint_kernel, 2)
        } else {
            if (first_mode == 0) TB_FP_CASE(tb_fos_kernel, 0) else if (first_mode == 1) TB_FP_CASE(tb_fos_kernel, 1)
            else TB_FP_CASE(tb_fos_kernel, 2)
        }
#undef TB_FP_CASE
        }
        TB_TRY(tb_check(ctx, cudaGetLastError(), "tb_footprint_kernel"));
    } else {
        TB_TRY(tb_tiles(ctx, S.ext_off, TB_ROLL_TILE, ctx->reg_e, n_tiles));
        tb_sc_view V{S.d_ext_off.as<i64>(), ctx->reg_d.as<i64>(), ctx->reg_e.as<i64>(), S.n_regions, p->flank};
        int w = p->kind == TB_SCORE_NONE ? 0 : p->window;
        TB_TRY(tb_launch_rolling<0>(ctx, V, n_tiles, S.signal.p, prep, w, p->kind == TB_SCORE_MEAN, first_mode, first_out));
    }
    if (smooth) {
        TB_TRY(tb_tiles(ctx, S.ext_off, TB_ROLL_TILE, ctx->reg_e, n_tiles));
        tb_sc_view V{S.d_ext_off.as<i64>(), ctx->reg_d.as<i64>(), ctx->reg_e.as<i64>(), S.n_regions, p->flank};
        TB_TRY(tb_launch_rolling<1>(ctx, V, n_tiles, S.work0.p, prep, p->smooth, 1, p->as_f64 ? 1 : 0, S.out.p));
    }
    tr.mark("kernels enqueued");
    tb_timer_stop(ctx);
    if (sink && !sink_served) {                                  // one copy behind the whole run
        TB_CUDA(ctx, cudaEventRecord(ctx->ev_chunk[0], ctx->stream));
        TB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_chunk[0], 0));
        TB_CUDA(ctx, cudaMemcpyAsync(sink, S.out.p, (size_t)S.out_total * (p->as_f64 ? 8 : 4), cudaMemcpyDeviceToHost, ctx->copy_stream));
    }
    TB_TRY(tb_sync(ctx, "tb_score_run"));
    S.have_out = true;
    return TB_OK;
}

extern "C" int tb_score_set_sink(tb_ctx *ctx, void *out) {
    if (!ctx) return TB_ERR_ARG;
    ctx->sc.sink = out;
    return TB_OK;
}

extern "C" int tb_score_sink_wait(tb_ctx *ctx) {
    if (!ctx) return TB_ERR_ARG;
    TB_TRY(tb_check(ctx, cudaStreamSynchronize(ctx->copy_stream), "tb_score_sink_wait"));
    return tb_check(ctx, cudaGetLastError(), "tb_score_sink_wait");
}

extern "C" int tb_score_download(tb_ctx *ctx, void *out) {
    if (!ctx || !out) return TB_ERR_ARG;
    auto &S = ctx->sc;
    if (!S.have_out) return tb_fail(ctx, TB_ERR_STATE, "tb_score_download: no completed tb_score_run");
    return tb_d2h(ctx, out, S.out.p, (size_t)S.out_total * (S.out_f64 ? 8 : 4));
}

// Scores of the last tb_score_run at chosen positions of the concatenated output space, without bringing the track to the host:
// what BINDetect's scan_and_score reads from the footprint bigWig (tobias/tools/bindetect_functions.py:296-330: the score at the
// middle of every TFBS and at the random background positions of every region).
template <class T>
__global__ void tb_score_gather_kernel(const T *__restrict__ track, i64 total, const i64 *__restrict__ index, i64 n, double *__restrict__ out,
                                       int *__restrict__ bad) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const i64 q = index[i];
        if (q < 0 || q >= total) {
            out[i] = (double)NAN;
            *bad = 1;
        } else {
            out[i] = (double)track[q];
        }
    }
}

extern "C" int tb_score_gather(tb_ctx *ctx, int64_t n, const int64_t *index, double *out) {
    if (!ctx || n < 0 || (n > 0 && (!index || !out))) return TB_ERR_ARG;
    auto &S = ctx->sc;
    if (!S.have_out) return tb_fail(ctx, TB_ERR_STATE, "tb_score_gather: no completed tb_score_run");
    if (n == 0) return TB_OK;
    TB_TRY(tb_h2d(ctx, ctx->reg_a, index, sizeof(i64) * (size_t)n));
    TB_TRY(tb_reserve(ctx, ctx->scratch0, sizeof(double) * (size_t)n + 16));
    double *d_out = ctx->scratch0.as<double>();
    int *d_bad = (int *)(d_out + n);
    TB_CUDA(ctx, cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
    const i64 want = tb_div_up(n, 256), capg = (i64)ctx->sm_count * 16;
    const unsigned grid = (unsigned)(want < capg ? want : capg);
    if (S.out_f64) {
        auto kg = tb_score_gather_kernel<double>;
        TB_LAUNCH(ctx, kg, grid, 256, 0, (const double *)S.out.as<double>(), S.out_total, (const i64 *)ctx->reg_a.as<i64>(), (i64)n, d_out, d_bad);
    } else {
        auto kg = tb_score_gather_kernel<float>;
        TB_LAUNCH(ctx, kg, grid, 256, 0, (const float *)S.out.as<float>(), S.out_total, (const i64 *)ctx->reg_a.as<i64>(), (i64)n, d_out, d_bad);
    }
    TB_TRY(tb_check(ctx, cudaGetLastError(), "tb_score_gather_kernel"));
    TB_TRY(tb_d2h(ctx, out, d_out, sizeof(double) * (size_t)n));
    int bad = 0;
    TB_TRY(tb_d2h(ctx, &bad, d_bad, sizeof(int)));
    if (bad) return tb_fail(ctx, TB_ERR_ARG, "tb_score_gather: an index lies outside [0, %lld)", (long long)S.out_total);
    return TB_OK;
}

extern "C" int tb_score_regions(tb_ctx *ctx, const float *signal, int64_t n_regions, const int64_t *ext_len, const tb_score_params *p,
                                void *out) {
    TB_TRY(tb_signal_upload(ctx, signal, n_regions, ext_len));
    TB_TRY(tb_score_run(ctx, p));
    return tb_score_download(ctx, out);
}

extern "C" int tb_rolling(tb_ctx *ctx, const double *arr, int64_t n, int w, int is_mean, double *out) {
    if (!ctx || n < 0 || (n > 0 && (!arr || !out))) return TB_ERR_ARG;
    if (w < 1) return tb_fail(ctx, TB_ERR_ARG, "tb_rolling: w must be >= 1");
    if (n == 0) return TB_OK;
    std::vector<i64> off = {0, (i64)n};
    TB_TRY(tb_h2d(ctx, ctx->reg_a, off.data(), sizeof(i64) * 2));
    TB_TRY(tb_h2d(ctx, ctx->scratch0, arr, (size_t)n * sizeof(double)));
    TB_TRY(tb_reserve(ctx, ctx->scratch1, (size_t)n * sizeof(double)));
    i64 n_tiles = 0;
    TB_TRY(tb_tiles(ctx, off, TB_ROLL_TILE, ctx->reg_e, n_tiles));
    tb_sc_view V{ctx->reg_a.as<i64>(), ctx->reg_a.as<i64>(), ctx->reg_e.as<i64>(), 1, 0};
    tb_prep prep{0, 0, 0, 0.0, 0.0};
    tb_timer_start(ctx);
    TB_TRY(tb_launch_rolling<1>(ctx, V, n_tiles, ctx->scratch0.p, prep, w, is_mean, 1, ctx->scratch1.p));
    tb_timer_stop(ctx);
    return tb_d2h(ctx, out, ctx->scratch1.p, (size_t)n * sizeof(double));
}

// ------------------------------------------------------------------------------------------------ gathers of ScoreBed / PlotAggregate
// Per-region summaries of the uploaded signal (tb_signal_upload: one extended region = one BED region; NaN -> 0 as
// OneRegion.get_signal does, tobias/utils/regions.py:172): what ScoreBed's score functions compute (tobias/tools/score_bed.py:30-54).
// op 0 start, 1 mid (index int(len / 2)), 2 end, 3 min, 4 max, 5 mean, 6 sum.  One warp per region, fp64 accumulation.
__global__ void tb_signal_summary_kernel(const float *__restrict__ sig, const i64 *__restrict__ off, i64 n_regions, int op, double *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((i64)gridDim.x * blockDim.x) >> 5;
    for (i64 r = warp; r < n_regions; r += n_warps) {
        const i64 a = off[r], b = off[r + 1], len = b - a;
        auto val = [&](i64 i) { float v = sig[a + i]; return (v != v) ? 0.0 : (double)(v > FLT_MAX ? FLT_MAX : (v < -FLT_MAX ? -FLT_MAX : v)); };
        double res = 0.0;
        if (op <= 2) {
            res = val(op == 0 ? 0 : (op == 1 ? (i64)((double)len / 2.0) : len - 1));
        } else {
            double acc = op == 3 ? INFINITY : (op == 4 ? -INFINITY : 0.0);
            for (i64 i = lane; i < len; i += 32) {
                const double v = val(i);
                acc = op == 3 ? fmin(acc, v) : (op == 4 ? fmax(acc, v) : acc + v);
            }
            for (int d = 16; d > 0; d >>= 1) {
                const double o = __shfl_xor_sync(0xffffffffu, acc, d);
                acc = op == 3 ? fmin(acc, o) : (op == 4 ? fmax(acc, o) : acc + o);
            }
            res = op == 5 ? acc / (double)len : acc;
        }
        if (lane == 0) out[r] = res;
    }
}

extern "C" int tb_signal_summary(tb_ctx *ctx, int op, double *out) {
    if (!ctx || !out) return TB_ERR_ARG;
    auto &S = ctx->sc;
    if (!S.have_signal) return tb_fail(ctx, TB_ERR_STATE, "tb_signal_summary: no signal (tb_signal_upload)");
    if (op < 0 || op > 6) return tb_fail(ctx, TB_ERR_ARG, "tb_signal_summary: op must be in [0, 6]");
    if (S.n_regions == 0) return TB_OK;
    TB_TRY(tb_reserve(ctx, ctx->scratch0, sizeof(double) * (size_t)S.n_regions));
    const i64 want = tb_div_up(S.n_regions * 32, 256), capg = (i64)ctx->sm_count * 16;
    TB_LAUNCH(ctx, tb_signal_summary_kernel, (unsigned)(want < capg ? want : capg), 256, 0, (const float *)S.signal.as<float>(),
              (const i64 *)S.d_ext_off.as<i64>(), S.n_regions, op, ctx->scratch0.as<double>());
    TB_TRY(tb_check(ctx, cudaGetLastError(), "tb_signal_summary_kernel"));
    return tb_d2h(ctx, out, ctx->scratch0.p, sizeof(double) * (size_t)S.n_regions);
}

// Column means over the rows of a signal matrix = the uploaded regions, all of one width (PlotAggregate, tobias/tools/plot_aggregate.py:
// 236-257): rows with keep[r] == 0 are left out (outlier removal), rows with reverse[r] != 0 are read back to front (OneRegion.get_signal on
// the minus strand, regions.py:174-175), log_transform: sign(v) * log2(|v| + 1).  One thread per column, fp64 sums.
__global__ void tb_signal_aggregate_kernel(const float *__restrict__ sig, i64 n_rows, int width, const uint8_t *__restrict__ reverse,
                                           const uint8_t *__restrict__ keep, int log_transform, double *__restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= width) return;
    double acc = 0.0;
    i64 n = 0;
    for (i64 r = 0; r < n_rows; r++) {
        if (keep && !keep[r]) continue;
        float v = sig[r * width + ((reverse && reverse[r]) ? width - 1 - c : c)];
        double d = (v != v) ? 0.0 : (double)v;
        if (log_transform) d = d < 0.0 ? -log2(-d + 1.0) : log2(d + 1.0);
        acc += d;
        n++;
    }
    out[c] = n > 0 ? acc / (double)n : (double)NAN;
}

extern "C" int tb_signal_aggregate(tb_ctx *ctx, const uint8_t *reverse, const uint8_t *keep, int log_transform, double *out) {
    if (!ctx || !out) return TB_ERR_ARG;
    auto &S = ctx->sc;
    if (!S.have_signal || S.n_regions == 0) return tb_fail(ctx, TB_ERR_STATE, "tb_signal_aggregate: no signal (tb_signal_upload)");
    const i64 width = S.ext_off[1] - S.ext_off[0];
    for (i64 r = 0; r < S.n_regions; r++)
        if (S.ext_off[r + 1] - S.ext_off[r] != width) return tb_fail(ctx, TB_ERR_ARG, "tb_signal_aggregate: regions differ in width (%lld)", (long long)r);
    if (width > (1 << 24)) return tb_fail(ctx, TB_ERR_ARG, "tb_signal_aggregate: width too large");
    const uint8_t *d_rev = nullptr, *d_keep = nullptr;
    if (reverse) {
        TB_TRY(tb_h2d(ctx, ctx->reg_a, reverse, (size_t)S.n_regions));
        d_rev = ctx->reg_a.as<uint8_t>();
    }
    if (keep) {
        TB_TRY(tb_h2d(ctx, ctx->reg_b, keep, (size_t)S.n_regions));
        d_keep = ctx->reg_b.as<uint8_t>();
    }
    TB_TRY(tb_reserve(ctx, ctx->scratch0, sizeof(double) * (size_t)width));
    TB_LAUNCH(ctx, tb_signal_aggregate_kernel, (unsigned)tb_div_up(width, 128), 128, 0, (const float *)S.signal.as<float>(), S.n_regions, (int)width,
              d_rev, d_keep, log_transform, ctx->scratch0.as<double>());
    TB_TRY(tb_check(ctx, cudaGetLastError(), "tb_signal_aggregate_kernel"));
    return tb_d2h(ctx, out, ctx->scratch0.p, sizeof(double) * (size_t)width);
}
