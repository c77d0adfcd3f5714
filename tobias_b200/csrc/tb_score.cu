// tb_score.cu -- K3a: per-position Tn5 bias score (DWM / PWM) over 2-bit packed genome sequence.
//
// Reference: DiNucleotideMatrix.score_sequence, tobias/utils/sequences.pyx:263-342 (PWM: :123-151), called per
// strand by bias_correction, tobias/tools/atacorrect_functions.py:254-262 (reverse strand = score of the reverse
// complement, reversed; NaN -> 0).
//
// Reformulation.  For a window with bases S_0..S_{L-1} the reference computes, per model x in {bias, background},
//     E_x[n][a] = PWM_x[a,n] + sum_{m != n} (DWM_x[S_m, a, m, n] - PWM_x[a,n])                (a = 0..3)
//     p_x[n]    = E_x[n][S_n] - log2( sum_a 2^E_x[n][a] ),      score = sum_n p_bias[n] - sum_n p_bg[n].
// With U_x[m][s][n][a] = DWM_x[s,a,m,n] - PWM_x[a,n] (m != n), U_x[n][s][n][a] = PWM_x[a,n], this is
//     E_x[n][.] = sum_m U_x[m][S_m][n][.]
// i.e. 25 selected rows of a 100 x 100 table are summed.  The table of BOTH models of one strand (160 KB of fp64 for
// L = 25) is staged once per block in shared memory and stays resident while the block streams positions
// (persistent blocks, one per SM, half of them per strand).  Within a warp lane n owns column n: its four
// accumulators come from two 16-byte halves stored [half][n] so every row read is two conflict-free LDS.128, and S_m is warp-
// uniform.  The reverse-strand table is pre-permuted on the host (m -> L-1-m, n -> L-1-n, bases complemented) so
// that both strands run the same code on forward-oriented k-mers.  All arithmetic is fp64.
#include "tb_common.cuh"

struct tb_score_regions_view {
    const i64 *ext_start;   // linear start of each (extended) region
    const i64 *ext_off;     // offsets into the concatenated position space, n+1 entries
    i64 n;
    i64 total;
};

#define TB_K3A_WPI 2        // windows in flight per warp (independent accumulator sets)

__global__ void __launch_bounds__(768, 1) tb_score_dwm_kernel(tb_genome_view g, tb_score_regions_view R, const double *__restrict__ table_fwd,
                                    const double *__restrict__ table_rev, int L, int use_smem, int strand_mask, int nan_invalid,
                                    double *__restrict__ out_fwd, double *__restrict__ out_rev) {
    TB_DYN_SMEM(smem);
    const int k_flank = L / 2;
    const int tbl_doubles = 2 * L * 4 * L * 4;
    // strands handled by this block: both masks set -> even blocks forward, odd blocks reverse
    int strand, part, nparts;
    if (strand_mask == 3) {
        strand = blockIdx.x & 1;
        part = blockIdx.x >> 1;
        nparts = (gridDim.x + 1 - strand) >> 1;
    } else {
        strand = strand_mask == 2 ? 1 : 0;
        part = blockIdx.x;
        nparts = gridDim.x;
    }
    const double *src = strand ? table_rev : table_fwd;
    const double *U = src;
    if (use_smem) {
        double *dst = (double *)smem;
        for (int i = threadIdx.x; i < tbl_doubles; i += blockDim.x) dst[i] = src[i];
        U = dst;
        __syncthreads();
    }
    double *out = strand ? out_rev : out_fwd;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int warps_per_block = blockDim.x >> 5;
    const i64 n_units = (R.total + 31) >> 5;              // a unit = 32 consecutive positions
    const int row = L * 4;                                // doubles per (m, s) row
    const int nl = lane < L ? lane : L - 1;               // lanes >= L shadow lane L-1 (their result is discarded)
    const double invalid = nan_invalid ? (double)NAN : 0.0;

    for (i64 unit = (i64)part * warps_per_block + warp; unit < n_units; unit += (i64)nparts * warps_per_block) {
        i64 pos0 = unit << 5;
        i64 j = tb_upper_bound(R.ext_off, R.n + 1, pos0) - 1;
        for (int it = 0; it < 32; it += TB_K3A_WPI) {
            u64 kmer[TB_K3A_WPI];
            bool ok[TB_K3A_WPI];
            i64 pos[TB_K3A_WPI];
#pragma unroll
            for (int w = 0; w < TB_K3A_WPI; w++) {
                pos[w] = pos0 + it + w;
                ok[w] = false;
                kmer[w] = 0;
                if (pos[w] < R.total) {
                    while (pos[w] >= R.ext_off[j + 1]) j++;
                    i64 local = pos[w] - R.ext_off[j];
                    i64 len = R.ext_off[j + 1] - R.ext_off[j];
                    if (local >= k_flank && local < len - k_flank) {
                        i64 gp = R.ext_start[j] + local;
                        if (tb_nflags(g, gp - k_flank, L) == 0) {
                            kmer[w] = tb_bases(g, gp - k_flank, L);
                            ok[w] = true;
                        }
                    }
                }
            }
            double acc[TB_K3A_WPI][2][4];
#pragma unroll
            for (int w = 0; w < TB_K3A_WPI; w++)
#pragma unroll
                for (int x = 0; x < 2; x++)
#pragma unroll
                    for (int a = 0; a < 4; a++) acc[w][x][a] = 0.0;
            for (int m = 0; m < L; m++) {
#pragma unroll
                for (int w = 0; w < TB_K3A_WPI; w++) {
                    int s = (int)((kmer[w] >> (2 * m)) & 3u);
#pragma unroll
                    for (int x = 0; x < 2; x++) {
                        // row layout [half][n][2]: lane n reads 16 B at n*16 -> conflict-free LDS.128 (a 32-B stride would 2-way conflict)
                        const double2 *p = (const double2 *)(U + ((size_t)(x * L + m) * 4 + s) * row) + nl;
                        double2 v0 = p[0], v1 = p[L];
                        acc[w][x][0] += v0.x;
                        acc[w][x][1] += v0.y;
                        acc[w][x][2] += v1.x;
                        acc[w][x][3] += v1.y;
                    }
                }
            }
#pragma unroll
            for (int w = 0; w < TB_K3A_WPI; w++) {
                int sn = (int)((kmer[w] >> (2 * nl)) & 3u);
                double p[2];
#pragma unroll
                for (int x = 0; x < 2; x++) {
                    double cn = exp2(acc[w][x][0]);
                    cn += exp2(acc[w][x][1]);
                    cn += exp2(acc[w][x][2]);
                    cn += exp2(acc[w][x][3]);
                    double e = sn == 0 ? acc[w][x][0] : sn == 1 ? acc[w][x][1] : sn == 2 ? acc[w][x][2] : acc[w][x][3];
                    p[x] = e - log2(cn);
                }
                double pb = lane < L ? p[0] : 0.0;       // x = 0: bias model, x = 1: background model
                double pg = lane < L ? p[1] : 0.0;
                for (int d = 16; d > 0; d >>= 1) {
                    pb += __shfl_xor_sync(0xffffffffu, pb, d);
                    pg += __shfl_xor_sync(0xffffffffu, pg, d);
                }
                if (lane == 0 && pos[w] < R.total) out[pos[w]] = ok[w] ? pb - pg : invalid;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ K3a, product form
// The scoring kernel above is bound by shared-memory bandwidth (25 row reads of 800 B per model and window) and pays
// 100 exp2 + 25 log2 per model and window.  Second formulation, same mathematics:
//   * rows are grouped: PAIRS of window offsets, T[p][d] = U[2p][d & 3] + U[2p+1][d >> 2] for the 16 dinucleotides d, then
//     TRIPLES (64 trinucleotide rows each), then at most one single offset (4 rows); tb_k3a2_groups picks the grouping with
//     the fewest shared-memory groups that fits the SM: 5 pair + 5 triple groups = 10 row reads instead of 25 for L = 25;
//   * the table holds powers of two, so 2^E[n][a] is a product of 10 table entries -- no exp2 at all -- and the per-model
//     score  sum_n (E[n][S_n] - log2 sum_a 2^E[n][a])  =  log2 prod_n (2^E[n][S_n] / Z[n])  needs two log2 per window (of the
//     products of the 25 selected terms and of the 25 normalisers, exponents carried apart);
//   * ratio form (below): three values per column instead of four, lane n owns column n;
//   * the pair groups are served from tensor memory (below), the triples from shared memory.
// One model of one strand is resident per block; the four (strand, model) classes of blocks write partial scores that
// tb_score_combine_kernel subtracts (bias - background).  A warp owns 32 consecutive positions: lane w prepares window
// w (region lookup, N test, packed k-mer) and finishes it (log2, one coalesced store of the 32 results); in between the
// warp walks the windows four at a time.  The cross-lane products of the four windows are reduced together (each butterfly
// level halves the number of windows a lane still carries), 8.75 shuffles per window.  Relative error of a product of 10
// correctly rounded factors ~ 2e-15, i.e. ~1e-14 absolute on the score -- the same order as the sum form's difference to
// the reference's libm.
#ifndef TB_K3A2_THREADS
#define TB_K3A2_THREADS 768
#endif
#define TB_K3A2_WPI 4

__device__ __forceinline__ void tb_split_exp(double v, double &mant, int &ex) {      // v = mant * 2^ex, mant in [1, 2)  (v > 0, normal)
    int hi = __double2hiint(v);
    ex = ((hi >> 20) & 0x7ff) - 1023;
    mant = __hiloint2double((hi & 0x800fffff) | 0x3ff00000, __double2loint(v));
}

// ---- ratio form of the rows -----------------------------------------------------------------------------------------
// The score of a column, 2^E[n][S_n] / sum_a 2^E[n][a], does not change when the four entries (a = 0..3) of a table row and
// column are scaled together, so every (row, n) quadruple is stored divided by its a = 0 entry: three ratios per column, the
// fourth is 1.  A row is L x 3 doubles (600 B for L = 25, 25 % less shared-memory traffic and one product less per column), and
// with lane n owning column n the normaliser Z[n] = 1 + r1 + r2 + r3 and the selected term need no cross-lane work at all.
// Row layout: (r1, r2) pairs of the L columns, then the L values r3, padded to a multiple of 16 bytes.
#define TB_K3A2_TMEM_MAX 5                                    // pair groups in tensor memory: 6 columns per row, 16 rows per group
__host__ __device__ static inline int tb_k3a2_rowstride(int L) { return (3 * L + 1) & ~1; }   // doubles

// Row groups of the product-form table for model length L: np dinucleotide groups (offsets 2p, 2p+1; 16 rows each), then nt3
// trinucleotide groups (64 rows each), then ns (0 / 1) single offsets (4 rows).  The first ntm pair groups live in tensor
// memory; the rest must fit the shared memory of one SM.  Fewest shared-memory groups wins (= fewest shared-memory row reads).
static inline void tb_k3a2_groups(int L, int tmem_groups, int &np, int &nt3, int &ns, int &ntm) {
    const int budget = (int)(((size_t)227 * 1024) / ((size_t)tb_k3a2_rowstride(L) * sizeof(double)));
    int best = 1 << 30;
    np = L / 2; nt3 = 0; ns = L & 1; ntm = 0;
    for (int t3 = L / 3; t3 >= 0; t3--) {
        const int rem = L - 3 * t3, p = rem / 2, s1 = rem & 1;
        const int tm = p < tmem_groups ? p : tmem_groups;
        const int rows_smem = 16 * (p - tm) + 64 * t3 + 4 * s1;
        const int groups = (p - tm) + t3 + s1;
        if (rows_smem <= budget && groups < best) {
            best = groups;
            np = p; nt3 = t3; ns = s1; ntm = tm;
        }
    }
}
static inline int tb_k3a2_tmem_groups(const tb_ctx *ctx) {
#ifdef TB_EMU
    (void)ctx;
    return 0;
#else
    return ctx->k3a_no_tmem ? 0 : (ctx->k3a_tmem_groups > 0 && ctx->k3a_tmem_groups <= TB_K3A2_TMEM_MAX ? ctx->k3a_tmem_groups : TB_K3A2_TMEM_MAX);
#endif
}
static inline bool tb_k3a2_fits(int L, int tmem_groups) {
    int np, nt3, ns, ntm;
    tb_k3a2_groups(L, tmem_groups, np, nt3, ns, ntm);
    return (size_t)(16 * (np - ntm) + 64 * nt3 + 4 * ns) * tb_k3a2_rowstride(L) * sizeof(double) <= (size_t)227 * 1024;
}

// Tensor memory (TMEM, 256 KB per SM on sm_100) as a second home for table rows: it is read through its own datapath
// (tcgen05.ld), so rows served from it cost no shared-memory wavefronts.  512 columns x 128 lanes x 32 bit: a warp reaches the
// 32 lanes of its quarter (warp id mod 4), so the rows are replicated in the four quarters; lane n of a quarter holds the three
// ratios of column n of a row in 6 consecutive columns -> 80 rows = the first five dinucleotide groups.
#ifndef TB_EMU
__device__ __forceinline__ void tb_tmem_alloc512(u32 *smem_slot) {       // one warp; writes the base address to shared memory
    const u32 a = (u32)__cvta_generic_to_shared(smem_slot);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(a) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tb_tmem_dealloc512(u32 taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(taddr) : "memory");
}
__device__ __forceinline__ void tb_tmem_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tb_tmem_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tb_tmem_st6(u32 taddr, double r1, double r2, double r3) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(taddr), "r"(__double2loint(r1)),
                 "r"(__double2hiint(r1)), "r"(__double2loint(r2)), "r"(__double2hiint(r2))
                 : "memory");
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};\n" ::"r"(taddr + 4u), "r"(__double2loint(r3)),
                 "r"(__double2hiint(r3))
                 : "memory");
}
__device__ __forceinline__ void tb_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tb_tmem_ld8(u32 taddr, u32 (&v)[8]) {      // 6 columns of this row + 2 of the next (ignored)
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tb_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
#endif

// FIXED: the grouping of the default model length (L = 25: 5 pair groups in tensor memory, 5 trinucleotide groups in shared
// memory) as compile-time constants, so the row loops unroll and the row offsets fold into the load instructions.
template <bool FIXED>
__global__ void __launch_bounds__(TB_K3A2_THREADS, 1)
    tb_score_dwm2_kernel(tb_genome_view g, tb_score_regions_view R, const double *__restrict__ vt_fwd, const double *__restrict__ vt_rev,
                         int L_, int np_, int nt3_, int ns_, int ntm_, int strand_mask, double *__restrict__ part) {
    TB_DYN_SMEM(smem);
    const unsigned FULL = 0xffffffffu;
    const int L = FIXED ? 25 : L_, np = FIXED ? 5 : np_, nt3 = FIXED ? 5 : nt3_, ns = FIXED ? 0 : ns_;
    int ntm = FIXED ? 5 : ntm_;
    const int k_flank = L / 2;
    const int rows = 16 * np + 64 * nt3 + 4 * ns, rs = tb_k3a2_rowstride(L);   // row stride in doubles; the first 16 ntm rows live in tensor memory
    const int nclass = strand_mask == 3 ? 4 : 2;
    const int cls = blockIdx.x % nclass;
    const int bpart = blockIdx.x / nclass, nparts = (gridDim.x - cls + nclass - 1) / nclass;
    const int strand = strand_mask == 3 ? (cls >> 1) : (strand_mask == 2 ? 1 : 0);
    const int model = cls & 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const bool col = lane < L;                                // lane n owns column n
    const double *src = (strand ? vt_rev : vt_fwd) + (size_t)model * rows * rs;
#ifndef TB_EMU
    __shared__ u32 tmem_slot;
    u32 tq = 0;                                               // this warp's quarter of the tensor memory
    if (ntm) {
        if (warp == 0) tb_tmem_alloc512(&tmem_slot);
        tb_tmem_fence_before();
        __syncthreads();
        tb_tmem_fence_after();
        tq = tmem_slot + ((u32)(32 * (warp & 3)) << 16);
        if (warp < 4) {                                       // warps 0..3 fill the four quarters
            for (int r = 0; r < 16 * ntm; r++) {
                const double *rp = src + (size_t)r * rs;
                const double r1 = col ? rp[2 * lane] : 1.0, r2 = col ? rp[2 * lane + 1] : 1.0, r3 = col ? rp[2 * L + lane] : 1.0;
                tb_tmem_st6(tq + 6u * (u32)r, r1, r2, r3);
            }
            tb_tmem_wait_st();
        }
        tb_tmem_fence_before();
    }
#else
    ntm = 0;
#endif
    {
        const double2 *s2 = (const double2 *)(src + (size_t)16 * ntm * rs);
        double2 *dst = (double2 *)smem;
        for (int i = threadIdx.x; i < (rows - 16 * ntm) * rs / 2; i += blockDim.x) dst[i] = s2[i];
        __syncthreads();
    }
#ifndef TB_EMU
    if (ntm) tb_tmem_fence_after();
#endif
    double *out = part + (size_t)(strand * 2 + model) * R.total;
    // row r >= 16 ntm: (r1, r2) of this lane's column at V12[r * rs], r3 at V3[r * rs]
    const int cl = col ? lane : 0;
    const double *V12 = (const double *)smem + 2 * cl - (size_t)16 * ntm * rs;
    const double *V3 = (const double *)smem + 2 * L + cl - (size_t)16 * ntm * rs;
    const int n_tail = nt3 + ns;                              // groups after the pairs: trinucleotides, then the single offset
    const int shcl = 2 * cl;
    const u32 snmask = col ? 3u : 0u;
    const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
    const int myw = 4 * (lane & 7) + ((lane >> 4) & 1) + 2 * ((lane >> 3) & 1);
    const i64 n_units = (R.total + 31) >> 5;

    for (i64 unit = (i64)bpart * wpb + warp; unit < n_units; unit += (i64)nparts * wpb) {
        const i64 pos0 = unit << 5;
        // ---- lane w prepares window w of the unit
        u32 klo = 0, khi = 0;
        bool okl = false;
        {
            const i64 pos = pos0 + lane;
            i64 j = tb_upper_bound_warp(R.ext_off, R.n + 1, pos0) - 1;     // region of the unit's first position, found by the whole warp
            if (pos < R.total) {
                while (pos >= R.ext_off[j + 1]) j++;                       // the lane's own region: the same or one of the next few
                i64 local = pos - R.ext_off[j];
                i64 len = R.ext_off[j + 1] - R.ext_off[j];
                if (local >= k_flank && local < len - k_flank) {
                    i64 gp = R.ext_start[j] + local;
                    if (tb_nflags(g, gp - k_flank, L) == 0) {
                        u64 km = tb_bases(g, gp - k_flank, L);
                        klo = (u32)km;
                        khi = (u32)(km >> 32);
                        okl = true;
                    }
                }
            }
        }
        const unsigned okmask = __ballot_sync(FULL, okl);
        double my_sel = 1.0, my_z = 1.0;
        int my_de = 0;
        for (int it = 0; it < 32; it += TB_K3A2_WPI) {
            if (((okmask >> it) & ((1u << TB_K3A2_WPI) - 1u)) == 0u) continue;       // warp-uniform
            u64 kmer[TB_K3A2_WPI];
            double a1[TB_K3A2_WPI], a2[TB_K3A2_WPI], a3[TB_K3A2_WPI];               // products of the ratios of column `lane`
#pragma unroll
            for (int w = 0; w < TB_K3A2_WPI; w++) {
                kmer[w] = (u64)__shfl_sync(FULL, klo, it + w) | ((u64)__shfl_sync(FULL, khi, it + w) << 32);
                a1[w] = a2[w] = a3[w] = 1.0;
            }
#ifndef TB_EMU
#pragma unroll
            for (int p = 0; p < (FIXED ? 5 : ntm); p++) {      // rows served by tensor memory
                u32 t[TB_K3A2_WPI][8];
#pragma unroll
                for (int w = 0; w < TB_K3A2_WPI; w++) {
                    const u32 d = (u32)(kmer[w] >> (4 * p)) & 15u;
                    tb_tmem_ld8(tq + 6u * (16u * (u32)p + d), t[w]);
                }
                tb_tmem_wait_ld();
#pragma unroll
                for (int w = 0; w < TB_K3A2_WPI; w++) {
                    a1[w] *= __hiloint2double((int)t[w][1], (int)t[w][0]);
                    a2[w] *= __hiloint2double((int)t[w][3], (int)t[w][2]);
                    a3[w] *= __hiloint2double((int)t[w][5], (int)t[w][4]);
                }
            }
#endif
#pragma unroll
            for (int p = (FIXED ? 5 : ntm); p < (FIXED ? 5 : np); p++) {
#pragma unroll
                for (int w = 0; w < TB_K3A2_WPI; w++) {
                    const int d = (int)((u32)(kmer[w] >> (4 * p)) & 15u);
                    const int ro = (p * 16 + d) * rs;
                    if (col) {
                        const double2 v = *(const double2 *)(V12 + ro);
                        a1[w] *= v.x;
                        a2[w] *= v.y;
                        a3[w] *= V3[ro];
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < (FIXED ? 5 : n_tail); t++) {
                const int sh = 4 * np + 6 * (t < nt3 ? t : nt3);
                const u32 mask = t < nt3 ? 63u : 3u;
                const int rb = 16 * np + 64 * (t < nt3 ? t : nt3);
#pragma unroll
                for (int w = 0; w < TB_K3A2_WPI; w++) {
                    const int d = (int)((u32)(kmer[w] >> sh) & mask);
                    const int ro = (rb + d) * rs;
                    if (col) {
                        const double2 v = *(const double2 *)(V12 + ro);
                        a1[w] *= v.x;
                        a2[w] *= v.y;
                        a3[w] *= V3[ro];
                    }
                }
            }
            // ---- column totals in the owning lane: Z = 1 + r1 + r2 + r3 (the a = 0 term is the unit), selected term by S_n.
            // Lanes >= L carry a1 = a2 = a3 = 1 and select the unit (snmask = 0): sel = 1, Z = 4 exactly, i.e. every idle lane
            // adds -2 to the exponent balance of a window and nothing to the mantissas; the +2 (32 - L) goes back in at the end.
            double sel[TB_K3A2_WPI], zz[TB_K3A2_WPI];
#pragma unroll
            for (int w = 0; w < TB_K3A2_WPI; w++) {
                const u32 sn = (u32)(kmer[w] >> shcl) & snmask;
                zz[w] = ((1.0 + a1[w]) + a2[w]) + a3[w];
                // (1, a1, a2, a3)[S_n] word by word with selects: a branch here diverges four ways inside every warp
                const bool b0 = (sn & 1u) != 0u, b1 = (sn & 2u) != 0u;
                const int lo_h = b0 ? __double2hiint(a1[w]) : 0x3ff00000, lo_l = b0 ? __double2loint(a1[w]) : 0;
                const int hi_h = b0 ? __double2hiint(a3[w]) : __double2hiint(a2[w]), hi_l = b0 ? __double2loint(a3[w]) : __double2loint(a2[w]);
                sel[w] = __hiloint2double(b1 ? hi_h : lo_h, b1 ? hi_l : lo_l);
            }
            // ---- products over the 32 lanes, four windows at once: after level 16 a lane carries two windows, after level 8 one.
            // Level 16 multiplies the raw terms (two factors of magnitude < 2^480 each, tb_set_bias_model checks the table for
            // that); the products are then split into mantissa in [1, 2) and integer exponent and only mantissas are multiplied.
            static_assert(TB_K3A2_WPI == 4, "the reduction below is written for four windows in flight");
            double ps[2], pz[2];
            int pe[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {          // windows (0,1) -> ps[0], windows (2,3) -> ps[1]
                const int wk = 2 * h, ws = 2 * h + 1;
                const double ks = up16 ? sel[ws] : sel[wk], kz = up16 ? zz[ws] : zz[wk];
                const double s2 = ks * __shfl_xor_sync(FULL, up16 ? sel[wk] : sel[ws], 16);
                const double z2 = kz * __shfl_xor_sync(FULL, up16 ? zz[wk] : zz[ws], 16);
                int es, ez;
                tb_split_exp(s2, ps[h], es);
                tb_split_exp(z2, pz[h], ez);
                pe[h] = es - ez;
            }
            double rs_ = (up8 ? ps[1] : ps[0]) * __shfl_xor_sync(FULL, up8 ? ps[0] : ps[1], 8);
            double rz = (up8 ? pz[1] : pz[0]) * __shfl_xor_sync(FULL, up8 ? pz[0] : pz[1], 8);
            int re = (up8 ? pe[1] : pe[0]) + __shfl_xor_sync(FULL, up8 ? pe[0] : pe[1], 8);
#pragma unroll
            for (int dlt = 4; dlt > 0; dlt >>= 1) {
                rs_ *= __shfl_xor_sync(FULL, rs_, dlt);
                rz *= __shfl_xor_sync(FULL, rz, dlt);
                re += __shfl_xor_sync(FULL, re, dlt);
            }
            // the 8 lanes with (bit 4, bit 3) = (b4, b3) now all hold window it + b4 + 2 * b3: lane it / 4 of each group keeps it
            if ((lane & 7) == (it >> 2)) {
                my_sel = rs_;
                my_z = rz;
                my_de = re;
            }
        }
        // lane (b4, b3, j) finishes window 4 j + b4 + 2 b3 of the unit
        if (pos0 + myw < R.total)
            out[pos0 + myw] = ((okmask >> myw) & 1u) ? (log2(my_sel) - log2(my_z)) + (double)(my_de + 2 * (32 - L)) : (double)NAN;
    }
#ifndef TB_EMU
    if (ntm) {
        tb_tmem_fence_before();
        __syncthreads();
        if (warp == 0) tb_tmem_dealloc512(tmem_slot);
    }
#endif
}

// score = partial(bias model) - partial(background model); NaN partials mark windows without a score
__global__ void tb_score_combine_kernel(const double *__restrict__ part, i64 total, int strand_mask, int nan_invalid,
                                        double *__restrict__ out_fwd, double *__restrict__ out_rev) {
    const double invalid = nan_invalid ? (double)NAN : 0.0;
    i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
        for (int strand = 0; strand < 2; strand++)
            if ((strand_mask >> strand) & 1) {
                double b = part[(size_t)(strand * 2) * total + i], gq = part[(size_t)(strand * 2 + 1) * total + i];
                (strand ? out_rev : out_fwd)[i] = (b != b || gq != gq) ? invalid : b - gq;
            }
}

// PWM: score = sum_m pssm[S_m][m]; one thread per position, table in shared memory.
__global__ void tb_score_pwm_kernel(tb_genome_view g, tb_score_regions_view R, const double *__restrict__ pssm_fwd,
                                    const double *__restrict__ pssm_rev, int L, int strand_mask, int nan_invalid,
                                    double *__restrict__ out_fwd, double *__restrict__ out_rev) {
    TB_DYN_SMEM(smem);
    double *T = (double *)smem;                  // [2 strands][L][4], reverse strand already permuted to forward orientation
    for (int i = threadIdx.x; i < 2 * L * 4; i += blockDim.x) {
        int strand = i / (L * 4), r = i % (L * 4), m = r / 4, s = r % 4;
        T[i] = strand == 0 ? pssm_fwd[s * L + m] : pssm_rev[(s ^ 1) * L + (L - 1 - m)];
    }
    __syncthreads();
    const int k_flank = L / 2;
    const double invalid = nan_invalid ? (double)NAN : 0.0;
    i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 pos = (i64)blockIdx.x * blockDim.x + threadIdx.x; pos < R.total; pos += stride) {
        i64 j = tb_upper_bound(R.ext_off, R.n + 1, pos) - 1;
        i64 local = pos - R.ext_off[j], len = R.ext_off[j + 1] - R.ext_off[j];
        bool ok = local >= k_flank && local < len - k_flank;
        u64 kmer = 0;
        if (ok) {
            i64 gp = R.ext_start[j] + local;
            ok = tb_nflags(g, gp - k_flank, L) == 0;
            if (ok) kmer = tb_bases(g, gp - k_flank, L);
        }
        for (int strand = 0; strand < 2; strand++) {
            if (!((strand_mask >> strand) & 1)) continue;
            double sc = invalid;
            if (ok) {
                sc = 0.0;
                if (strand == 0) {
                    for (int m = 0; m < L; m++) sc += T[m * 4 + (int)((kmer >> (2 * m)) & 3u)];
                } else {        // read-oriented summation order: m' = 0..L-1  <->  forward offset L-1-m'
                    for (int m = L - 1; m >= 0; m--) sc += T[L * 4 + m * 4 + (int)((kmer >> (2 * m)) & 3u)];
                }
            }
            (strand ? out_rev : out_fwd)[pos] = sc;
        }
    }
}

// ---------------------------------------------------------------------------------------------- host side
extern "C" int tb_set_bias_model(tb_ctx *ctx, int strand, int is_dwm, int L, const double *bias_pwm_log,
                                 const double *bg_pwm_log, const double *bias_dwm_log, const double *bg_dwm_log) {
    if (!ctx) return TB_ERR_ARG;
    if (strand < 0 || strand > 1) return tb_fail(ctx, TB_ERR_ARG, "tb_set_bias_model: strand must be 0 or 1");
    if (L < 1 || L > 31 || !(L & 1)) return tb_fail(ctx, TB_ERR_ARG, "tb_set_bias_model: L must be odd and <= 31");
    if (!bias_pwm_log || (is_dwm && (!bg_pwm_log || !bias_dwm_log || !bg_dwm_log)))
        return tb_fail(ctx, TB_ERR_ARG, "tb_set_bias_model: NULL table");
    tb_model &M = ctx->model[strand];
    M.is_dwm = is_dwm ? 1 : 0;
    M.L = L;
    ctx->cor.valid = false;
    if (!is_dwm) return tb_h2d(ctx, M.table, bias_pwm_log, sizeof(double) * 5 * L) || tb_sync(ctx, "tb_set_bias_model");
    // U[x][m][s][a>>1][n][a&1] (see the kernel for the layout); for the reverse strand the table is expressed on forward-oriented k-mers:
    //   U'[m'][s'][n'][a'] = U[L-1-m'][s'^1][L-1-n'][a'^1]
    std::vector<double> U((size_t)2 * L * 4 * L * 4);
    const double *pwm[2] = {bias_pwm_log, bg_pwm_log};
    const double *dwm[2] = {bias_dwm_log, bg_dwm_log};
    for (int x = 0; x < 2; x++)
        for (int m = 0; m < L; m++)
            for (int s = 0; s < 4; s++)
                for (int n = 0; n < L; n++)
                    for (int a = 0; a < 4; a++) {
                        int rm = strand ? L - 1 - m : m, rn = strand ? L - 1 - n : n;
                        int rs = strand ? s ^ 1 : s, ra = strand ? a ^ 1 : a;
                        double p = pwm[x][ra * L + rn];
                        double v = rm == rn ? p : dwm[x][(((size_t)rs * 5 + ra) * L + rm) * L + rn] - p;
                        U[(((((size_t)x * L + m) * 4 + s) * 2 + (a >> 1)) * L + n) * 2 + (a & 1)] = v;
                    }
    TB_TRY(tb_h2d(ctx, M.table, U.data(), U.size() * sizeof(double)));
    // product-form table (tb_score_dwm2_kernel), ratio form: with T[x][row][n][a] = the summed U rows of the group (pair group p: row
    // 16 p + d, offsets 2p (d & 3) and 2p + 1 (d >> 2); then the trinucleotide groups, 64 rows each; then a single offset, 4 rows),
    // a row stores 2^(T[a] - T[0]) for a = 1..3: (r1, r2) pairs of the L columns, then the L values r3; same forward-oriented
    // indexing as U
    {
        int np, nt3, ns, ntm;
        tb_k3a2_groups(L, tb_k3a2_tmem_groups(ctx), np, nt3, ns, ntm);
        const int rows = 16 * np + 64 * nt3 + 4 * ns, rs = tb_k3a2_rowstride(L);
        auto u = [&](int x, int m, int s_, int n, int a) {
            return U[(((((size_t)x * L + m) * 4 + s_) * 2 + (a >> 1)) * L + n) * 2 + (a & 1)];
        };
        std::vector<double> Vt((size_t)2 * rows * rs, 1.0);
        // range of a column product: sum over the groups of the largest |T[a] - T[0]| of the group (log2 units)
        const int ngroups = np + nt3 + ns;
        std::vector<double> span((size_t)2 * L * 3 * ngroups, 0.0);
        for (int x = 0; x < 2; x++)
            for (int r = 0; r < rows; r++)
                for (int n = 0; n < L; n++) {
                    double t[4];
                    for (int a = 0; a < 4; a++) {
                        if (r < 16 * np) {
                            int p = r >> 4, d = r & 15;
                            t[a] = u(x, 2 * p, d & 3, n, a) + u(x, 2 * p + 1, d >> 2, n, a);
                        } else if (r < 16 * np + 64 * nt3) {
                            int q = (r - 16 * np) >> 6, d = (r - 16 * np) & 63, m0 = 2 * np + 3 * q;
                            t[a] = (u(x, m0, d & 3, n, a) + u(x, m0 + 1, (d >> 2) & 3, n, a)) + u(x, m0 + 2, d >> 4, n, a);
                        } else {
                            t[a] = u(x, L - 1, r - 16 * np - 64 * nt3, n, a);
                        }
                    }
                    double *rp = Vt.data() + ((size_t)x * rows + r) * rs;
                    rp[2 * n] = exp2(t[1] - t[0]);
                    rp[2 * n + 1] = exp2(t[2] - t[0]);
                    rp[2 * L + n] = exp2(t[3] - t[0]);
                    const int grp = r < 16 * np ? r >> 4 : (r < 16 * np + 64 * nt3 ? np + ((r - 16 * np) >> 6) : np + nt3);
                    for (int a = 1; a < 4; a++) {
                        double &sp = span[(((size_t)x * L + n) * 3 + (a - 1)) * ngroups + grp];
                        const double d = fabs(t[a] - t[0]);
                        if (!(d <= sp)) sp = d;              // NaN / inf end up in the span as well
                    }
                }
        M.product_ok = 1;
        for (size_t c = 0; c < (size_t)2 * L * 3; c++) {
            double tot = 0.0;
            for (int q = 0; q < ngroups; q++) tot += span[c * ngroups + q];
            if (!(tot < 480.0)) M.product_ok = 0;           // tb_score_device falls back to the sum form (exp2 / log2 per column)
        }
        TB_TRY(tb_h2d(ctx, M.vtable, Vt.data(), Vt.size() * sizeof(double)));
    }
    return tb_sync(ctx, "tb_set_bias_model");
}

extern "C" int tb_bias_model_info(tb_ctx *ctx, int strand, int *is_dwm, int *L, int *product_form) {
    if (!ctx) return TB_ERR_ARG;
    if (strand < 0 || strand > 1) return tb_fail(ctx, TB_ERR_ARG, "tb_bias_model_info: strand must be 0 or 1");
    const tb_model &M = ctx->model[strand];
    if (is_dwm) *is_dwm = M.is_dwm;
    if (L) *L = M.L;
    if (product_form)
        *product_form = M.is_dwm == 1 && M.product_ok && !ctx->force_sum_form && tb_k3a2_fits(M.L, tb_k3a2_tmem_groups(ctx)) ? 1 : 0;
    return TB_OK;
}

// scores both strands (mask 3) or one; d_out_* are device arrays over the concatenated position space
int tb_score_device(tb_ctx *ctx, const i64 *d_ext_start, const i64 *d_ext_off, i64 n_regions, i64 total, int strand_mask,
                    int nan_invalid, double *d_out_fwd, double *d_out_rev) {
    int s0 = (strand_mask & 1) ? 0 : 1;
    tb_model &M0 = ctx->model[s0];
    for (int s = 0; s < 2; s++)
        if ((strand_mask >> s) & 1) {
            if (ctx->model[s].is_dwm < 0) return tb_fail(ctx, TB_ERR_STATE, "no bias model for strand %d: call tb_set_bias_model", s);
            if (ctx->model[s].is_dwm != M0.is_dwm || ctx->model[s].L != M0.L)
                return tb_fail(ctx, TB_ERR_STATE, "forward and reverse bias models differ in type or length");
        }
    if (total == 0) return TB_OK;
    TB_TRY(tb_genome_fence(ctx, 0, ctx->n_contigs - 1));
    tb_genome_view g{ctx->seq2.as<u32>(), ctx->nmask.as<u32>(), ctx->genome_bases};
    tb_score_regions_view R{d_ext_start, d_ext_off, n_regions, total};
    const double *tf = ctx->model[0].table.as<double>(), *tr = ctx->model[1].table.as<double>();
    if (!tf) tf = tr;
    if (!tr) tr = tf;
    int L = M0.L;
    int k3a_np, k3a_nt3, k3a_ns, k3a_ntm;
    tb_k3a2_groups(L, tb_k3a2_tmem_groups(ctx), k3a_np, k3a_nt3, k3a_ns, k3a_ntm);
    const size_t smem2 = (size_t)(16 * (k3a_np - k3a_ntm) + 64 * k3a_nt3 + 4 * k3a_ns) * tb_k3a2_rowstride(L) * sizeof(double);
    if (M0.is_dwm && tb_k3a2_fits(L, tb_k3a2_tmem_groups(ctx)) && !ctx->force_sum_form && ctx->model[0].product_ok && ctx->model[1].product_ok) {
        // product form: one (strand, model) table per block, partial scores, then bias - background
        const int nclass = strand_mask == 3 ? 4 : 2;
        TB_TRY(tb_reserve(ctx, ctx->scratch4, (size_t)4 * total * sizeof(double)));
        const bool fixed = L == 25 && k3a_np == 5 && k3a_nt3 == 5 && k3a_ns == 0 && k3a_ntm == 5;
        auto k = fixed ? tb_score_dwm2_kernel<true> : tb_score_dwm2_kernel<false>;
        TB_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        i64 n_units = (total + 31) >> 5;
        i64 want = tb_div_up(n_units, TB_K3A2_THREADS / 32) * nclass;
        i64 cap = ctx->sm_count - ctx->sm_count % nclass;
        unsigned grid = (unsigned)(want < cap ? want : cap);
        if (grid < (unsigned)nclass) grid = nclass;
        const double *vf = ctx->model[0].vtable.as<double>(), *vr = ctx->model[1].vtable.as<double>();
        if (!vf) vf = vr;
        if (!vr) vr = vf;
        TB_LAUNCH(ctx, k, grid, TB_K3A2_THREADS, smem2, g, R, vf, vr, L, k3a_np, k3a_nt3, k3a_ns, k3a_ntm, strand_mask, ctx->scratch4.as<double>());
        TB_TRY(tb_check(ctx, cudaGetLastError(), "tb_score_dwm2_kernel"));
        i64 wantc = tb_div_up(total, 256), capc = (i64)ctx->sm_count * 8;
        TB_LAUNCH(ctx, tb_score_combine_kernel, (unsigned)(wantc < capc ? wantc : capc), 256, 0,
                  (const double *)ctx->scratch4.as<double>(), total, strand_mask, nan_invalid, d_out_fwd, d_out_rev);
    } else if (M0.is_dwm) {
        size_t smem = (size_t)2 * L * 4 * L * 4 * sizeof(double);
        int use_smem = smem <= (size_t)200 * 1024;
        auto k = tb_score_dwm_kernel;
        if (use_smem) TB_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int block = 768;                  // 24 warps: 82 registers/thread fit once per SM next to the 160 KB table
        i64 n_units = (total + 31) >> 5;
        i64 want = tb_div_up(n_units, block / 32) * (strand_mask == 3 ? 2 : 1);
        i64 cap = ctx->sm_count;
        if (strand_mask == 3 && (cap & 1)) cap--;
        unsigned grid = (unsigned)(want < cap ? want : cap);
        if (strand_mask == 3 && grid < 2) grid = 2;
        TB_LAUNCH(ctx, k, grid, block, use_smem ? smem : 0, g, R, tf, tr, L, use_smem, strand_mask, nan_invalid, d_out_fwd, d_out_rev);
    } else {
        size_t smem = (size_t)2 * L * 4 * sizeof(double);
        i64 want = tb_div_up(total, 256), cap = (i64)ctx->sm_count * 8;
        TB_LAUNCH(ctx, tb_score_pwm_kernel, (unsigned)(want < cap ? want : cap), 256, smem, g, R, tf, tr, L, strand_mask, nan_invalid,
                  d_out_fwd, d_out_rev);
    }
    return tb_check(ctx, cudaGetLastError(), "tb_score kernel");
}

extern "C" int tb_score_sequence(tb_ctx *ctx, int strand, int contig, int64_t start, int64_t end, double *out) {
    if (!ctx || !out) return TB_ERR_ARG;
    if (strand < 0 || strand > 1) return tb_fail(ctx, TB_ERR_ARG, "tb_score_sequence: strand must be 0 or 1");
    std::vector<i64> ls, le;
    int32_t c = contig;
    int64_t s = start, e = end;
    TB_TRY(tb_linear_regions(ctx, 1, &c, &s, &e, ls, le, false));
    i64 total = le[0] - ls[0];
    i64 off[2] = {0, total};
    TB_TRY(tb_h2d(ctx, ctx->reg_a, ls.data(), sizeof(i64)));
    TB_TRY(tb_h2d(ctx, ctx->reg_c, off, sizeof(off)));
    TB_TRY(tb_reserve(ctx, ctx->scratch0, (size_t)total * sizeof(double)));
    tb_timer_start(ctx);
    TB_TRY(tb_score_device(ctx, ctx->reg_a.as<i64>(), ctx->reg_c.as<i64>(), 1, total, 1 << strand, 1, ctx->scratch0.as<double>(),
                           ctx->scratch0.as<double>()));
    tb_timer_stop(ctx);
    return tb_d2h(ctx, out, ctx->scratch0.p, (size_t)total * sizeof(double));
}
