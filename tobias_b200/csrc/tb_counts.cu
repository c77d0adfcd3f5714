// tb_counts.cu -- K2: cut-site weighted (di)nucleotide count tensors for the Tn5 bias model.
//
// Reference semantics: bias_estimation, tobias/tools/atacorrect_functions.py:109-181, and
// SequenceMatrix.add_sequence / add_background, tobias/utils/sequences.pyx:63-105 (PWM), :178-223 (DWM).
//   * a record belongs to every input region its alignment overlaps (fetch, :134)
//   * only records whose 5'-most CIGAR op is M count (:156-158)
//   * multiplicity n per (region, strand, cutsite), weight w = min(n, cap) (:159-168)
//   * the "within borders" test (:166) and the k-mer validity test (ngs.pyx:115) use the region extended by
//     k_flank + bg_shift and clipped to the contig (:142-143)  -- reference quirk, kept
//   * bias k-mer at the cut site, background k-mer bg_shift upstream of the read (:170-175)
//
// Two kernels.  (1) scatter: one thread per record, integer atomics into a per-(region, strand) multiplicity slab.
// (2) scan + accumulate: blocks stream the slab with coalesced loads, compact the non-zero sites into a shared-
// memory list of 2-bit packed k-mers, then every thread owns one (m, n) position pair and updates its private
// 16-entry (Sm, Sn) counter column in shared memory -- conflict-free, no atomics in the inner loop (the tensor is
// symmetric, so only m <= n is accumulated; the reverse strand is accumulated in forward orientation and
// reverse-complemented by an index permutation on the host).  Everything is integer: bit-exact in any order.
#include "tb_common.cuh"

struct tb_reads_view2 {
    const i64 *lstart;
    const int32_t *rlen;
    const int32_t *clip;
    const uint8_t *flags;
    i64 n;
};

struct tb_k2_regions {
    const i64 *s, *e;      // original bounds (linear)
    const i64 *es, *ee;    // extended + clipped bounds (linear)
    const i64 *soff;       // slab offset of each region (relative to the chunk), n+1 entries
    i64 n;                 // regions in this chunk
    i64 slab_len;          // entries per strand in this chunk
};

__global__ void tb_k2_scatter_kernel(tb_reads_view2 r, tb_k2_regions R, int shift_fwd, int shift_rev, u32 *__restrict__ slab) {
    i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < r.n; i += stride) {
        uint8_t f = r.flags[i];
        if ((f & TB_READ_SKIP) || !(f & TB_READ_5P_MATCH)) continue;
        int rev = (f & TB_READ_REVERSE) ? 1 : 0;
        i64 rs = r.lstart[i], re = rs + r.rlen[i];
        i64 p = rev ? re + r.clip[i] + shift_rev : rs - r.clip[i] + shift_fwd;
        for (i64 j = tb_upper_bound(R.e, R.n, rs); j < R.n && R.s[j] < re; j++) {
            if (p >= R.es[j] && p <= R.ee[j] - 2)                        // start < cutsite < end on the extended region
                atomicAdd(&slab[(i64)rev * R.slab_len + R.soff[j] + (p - R.es[j])], 1u);
        }
    }
}

#define TB_K2_TILE 1024

// layout of the accumulators: DWM: acc[tbl][code = Sm*4+Sn][pair]   PWM: acc[tbl][S (0..4)][m]
template <int IS_DWM>
__global__ void tb_k2_accumulate_kernel(tb_genome_view g, tb_k2_regions R, const u32 *__restrict__ slab, int k_flank,
                                        int bg_shift, int cap, i64 n_tiles, u64 *__restrict__ g_acc, u64 *__restrict__ g_scalars) {
    TB_DYN_SMEM(smem);
    const int L = 2 * k_flank + 1;
    const int npairs = L * (L + 1) / 2;
    const int n_acc = IS_DWM ? 4 * 16 * npairs : 4 * 5 * L;
    int *acc = (int *)smem;
    u64 *list = (u64 *)(smem + (((size_t)n_acc * 4 + 15) & ~(size_t)15));
    uint8_t *meta = (uint8_t *)(list + 4 * TB_K2_TILE);
    __shared__ int n_list;
    __shared__ int sc[5];            // no_reads, no_bias_f, no_bg_f, no_bias_r, no_bg_r
    const int tid = threadIdx.x;

    for (int i = tid; i < n_acc; i += blockDim.x) acc[i] = 0;
    if (tid < 5) sc[tid] = 0;
    // (m, n) pair owned by this thread (m <= n), row-major over the upper triangle
    int pm = 0, pn = 0;
    if (IS_DWM && tid < npairs) {
        int t = tid, m = 0;
        while (t >= L - m) { t -= L - m; m++; }
        pm = m;
        pn = m + t;
    }
    __syncthreads();

    for (i64 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (tid == 0) n_list = 0;
        __syncthreads();
        i64 x0 = tile * TB_K2_TILE;
        for (int y = tid; y < 2 * TB_K2_TILE; y += blockDim.x) {
            int strand = y / TB_K2_TILE;
            i64 x = x0 + (y - strand * TB_K2_TILE);
            if (x >= R.slab_len) continue;
            u32 cnt = slab[(i64)strand * R.slab_len + x];
            if (cnt == 0) continue;
            int w = (int)(cnt < (u32)cap ? cnt : (u32)cap);
            i64 j = tb_upper_bound(R.soff, R.n + 1, x) - 1;              // region owning slab slot x
            i64 es = R.es[j], ee = R.ee[j];
            i64 gp = es + (x - R.soff[j]);
            atomicAdd(&sc[0], w);
            for (int which = 0; which < 2; which++) {                    // 0: bias k-mer, 1: background k-mer
                i64 c = which == 0 ? gp : (strand ? gp + bg_shift : gp - bg_shift);
                bool valid = (c > es + k_flank) && (c <= ee - k_flank - 1);          // OneRead.get_kmer, ngs.pyx:115
                u64 kmer = 0;
                u32 nf = 0xffffffffu;
                if (valid) {
                    kmer = tb_bases(g, c - k_flank, L);
                    nf = tb_nflags(g, c - k_flank, L);
                }
                int tbl = strand * 2 + which;
                if (IS_DWM) {
                    if (valid && nf == 0) {
                        int slot = atomicAdd(&n_list, 1);
                        list[slot] = kmer;
                        meta[slot] = (uint8_t)(tbl | (w << 2));
                        atomicAdd(&sc[1 + tbl], w);
                    }
                } else {
                    // PWM: add_sequence skips k-mers holding N (sequences.pyx:73-76); add_background does not (:100-102)
                    if (which == 0 && !(valid && nf == 0)) continue;
                    for (int m = 0; m < L; m++) {
                        int mm = strand ? L - 1 - m : m;                 // position in the read-oriented k-mer
                        int s;
                        if (!valid || ((nf >> m) & 1u)) s = 4;
                        else {
                            s = (int)((kmer >> (2 * m)) & 3u);
                            if (strand) s ^= 1;
                        }
                        atomicAdd(&acc[(tbl * 5 + s) * L + mm], w);
                    }
                    atomicAdd(&sc[1 + tbl], w);
                }
            }
        }
        __syncthreads();
        if (IS_DWM && tid < npairs) {
            int nl = n_list;
            for (int e = 0; e < nl; e++) {
                u64 kmer = list[e];
                int mt = meta[e];
                int code = (int)(((kmer >> (2 * pm)) & 3u) * 4 + ((kmer >> (2 * pn)) & 3u));
                acc[((mt & 3) * 16 + code) * npairs + tid] += (mt >> 2);
            }
        }
        __syncthreads();
    }
    for (int i = tid; i < n_acc; i += blockDim.x)
        if (acc[i]) atomicAdd(&g_acc[i], (u64)(i64)acc[i]);
    if (tid < 5 && sc[tid]) atomicAdd(&g_scalars[tid], (u64)(i64)sc[tid]);
}

extern "C" int tb_bias_counts(tb_ctx *ctx, int64_t n_regions, const int32_t *contig, const int64_t *start, const int64_t *end,
                              int k_flank, int bg_shift, int shift_fwd, int shift_rev, int cap, int is_dwm, int64_t *counts_out,
                              int64_t *scalars_out) {
    if (!ctx || !counts_out || !scalars_out) return TB_ERR_ARG;
    if (k_flank < 0 || k_flank > 15) return tb_fail(ctx, TB_ERR_ARG, "tb_bias_counts: k_flank must be in [0, 15]");
    if (cap < 1 || cap > 63) return tb_fail(ctx, TB_ERR_ARG, "tb_bias_counts: cap must be in [1, 63]");
    if (bg_shift < 0) return tb_fail(ctx, TB_ERR_ARG, "tb_bias_counts: bg_shift must be >= 0");
    const int L = 2 * k_flank + 1;
    const int npairs = L * (L + 1) / 2;
    const size_t n_acc = is_dwm ? (size_t)4 * 16 * npairs : (size_t)4 * 5 * L;
    const size_t n_out = is_dwm ? (size_t)4 * 25 * L * L : (size_t)4 * 5 * L;
    memset(counts_out, 0, n_out * sizeof(int64_t));
    memset(scalars_out, 0, 5 * sizeof(int64_t));
    std::vector<i64> ls, le;
    TB_TRY(tb_linear_regions(ctx, n_regions, contig, start, end, ls, le, false));
    if (n_regions == 0 || ctx->n_reads == 0) return TB_OK;

    const i64 ext = (i64)k_flank + bg_shift;
    std::vector<i64> es(n_regions), ee(n_regions);
    for (i64 i = 0; i < n_regions; i++) {
        int c = contig[i];
        i64 lo = ctx->contig_off[c], hi = lo + ctx->contig_len[c];
        es[i] = ls[i] - ext < lo ? lo : ls[i] - ext;
        ee[i] = le[i] + ext > hi ? hi : le[i] + ext;
    }
    TB_TRY(tb_reserve(ctx, ctx->scratch1, (n_acc + 8) * sizeof(u64)));
    TB_CUDA(ctx, cudaMemsetAsync(ctx->scratch1.p, 0, (n_acc + 8) * sizeof(u64), ctx->stream));
    u64 *d_acc = ctx->scratch1.as<u64>();
    u64 *d_sc = d_acc + n_acc;

    tb_reads_view2 rv{ctx->r_start.as<i64>(), ctx->r_len.as<int32_t>(), ctx->r_clip.as<int32_t>(), ctx->r_flags.as<uint8_t>(),
                      ctx->n_reads};
    tb_genome_view g{ctx->seq2.as<u32>(), ctx->nmask.as<u32>(), ctx->genome_bases};
    const i64 SLAB_BUDGET = (i64)1 << 27;       // entries per strand per chunk (1 GiB of u32 for both strands)
    size_t smem = ((n_acc * 4 + 15) & ~(size_t)15) + (size_t)4 * TB_K2_TILE * 9;
    int block = is_dwm ? ((npairs + 31) / 32) * 32 : 256;
    if (block < 128) block = 128;
    if (is_dwm) {
        auto k = tb_k2_accumulate_kernel<1>;
        TB_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    } else {
        auto k = tb_k2_accumulate_kernel<0>;
        TB_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }

    tb_timer_start(ctx);
    i64 j0 = 0;
    std::vector<i64> soff;
    while (j0 < n_regions) {
        i64 j1 = j0, len = 0;
        soff.assign(1, 0);
        while (j1 < n_regions && (j1 == j0 || len + (ee[j1] - es[j1]) <= SLAB_BUDGET)) {
            len += ee[j1] - es[j1];
            soff.push_back(len);
            j1++;
        }
        i64 nr = j1 - j0;
        TB_TRY(tb_h2d(ctx, ctx->reg_a, ls.data() + j0, sizeof(i64) * nr));
        TB_TRY(tb_h2d(ctx, ctx->reg_b, le.data() + j0, sizeof(i64) * nr));
        TB_TRY(tb_h2d(ctx, ctx->reg_c, es.data() + j0, sizeof(i64) * nr));
        TB_TRY(tb_h2d(ctx, ctx->reg_d, ee.data() + j0, sizeof(i64) * nr));
        TB_TRY(tb_h2d(ctx, ctx->reg_e, soff.data(), sizeof(i64) * (nr + 1)));
        TB_TRY(tb_reserve(ctx, ctx->scratch0, (size_t)len * 2 * sizeof(u32)));
        TB_CUDA(ctx, cudaMemsetAsync(ctx->scratch0.p, 0, (size_t)len * 2 * sizeof(u32), ctx->stream));
        tb_k2_regions R{ctx->reg_a.as<i64>(), ctx->reg_b.as<i64>(), ctx->reg_c.as<i64>(), ctx->reg_d.as<i64>(),
                        ctx->reg_e.as<i64>(), nr, len};
        {
            i64 want = tb_div_up(rv.n, 256), capg = (i64)ctx->sm_count * 16;
            TB_LAUNCH(ctx, tb_k2_scatter_kernel, (unsigned)(want < capg ? want : capg), 256, 0, rv, R, shift_fwd, shift_rev,
                      ctx->scratch0.as<u32>());
        }
        i64 n_tiles = tb_div_up(len, TB_K2_TILE);
        i64 capg = (i64)ctx->sm_count * (is_dwm ? 1 : 2);
        unsigned grid = (unsigned)(n_tiles < capg ? n_tiles : capg);
        if (is_dwm) {
            auto k = tb_k2_accumulate_kernel<1>;
            TB_LAUNCH(ctx, k, grid, block, smem, g, R, (const u32 *)ctx->scratch0.as<u32>(), k_flank, bg_shift, cap, n_tiles, d_acc, d_sc);
        } else {
            auto k = tb_k2_accumulate_kernel<0>;
            TB_LAUNCH(ctx, k, grid, block, smem, g, R, (const u32 *)ctx->scratch0.as<u32>(), k_flank, bg_shift, cap, n_tiles, d_acc, d_sc);
        }
        TB_TRY(tb_sync(ctx, "tb_bias_counts chunk"));      // region tables are reused by the next chunk
        j0 = j1;
    }
    tb_timer_stop(ctx);

    std::vector<u64> h(n_acc + 8);
    TB_TRY(tb_d2h(ctx, h.data(), d_acc, (n_acc + 8) * sizeof(u64)));
    for (int i = 0; i < 5; i++) scalars_out[i] = (int64_t)h[n_acc + i];
    if (is_dwm) {
        // expand the upper triangle (forward orientation) into [strand][bias,bg][5][5][L][L]
        std::vector<int> pair_of((size_t)L * L);
        int t = 0;
        for (int m = 0; m < L; m++)
            for (int n = m; n < L; n++) pair_of[m * L + n] = t++;
        auto T = [&](int tbl, int a, int b, int m, int n) -> int64_t {      // full symmetric forward-oriented tensor
            if (m <= n) return (int64_t)h[((size_t)tbl * 16 + a * 4 + b) * npairs + pair_of[m * L + n]];
            return (int64_t)h[((size_t)tbl * 16 + b * 4 + a) * npairs + pair_of[n * L + m]];
        };
        for (int strand = 0; strand < 2; strand++)
            for (int which = 0; which < 2; which++) {
                int tbl = strand * 2 + which;
                int64_t *dst = counts_out + (size_t)tbl * 25 * L * L;
                for (int a = 0; a < 4; a++)
                    for (int b = 0; b < 4; b++)
                        for (int m = 0; m < L; m++)
                            for (int n = 0; n < L; n++)
                                dst[(((size_t)a * 5 + b) * L + m) * L + n] =
                                    strand == 0 ? T(tbl, a, b, m, n) : T(tbl, a ^ 1, b ^ 1, L - 1 - m, L - 1 - n);
            }
    } else {
        for (size_t i = 0; i < n_acc; i++) counts_out[i] = (int64_t)h[i];
    }
    return TB_OK;
}
