// tb_pileup.cu -- K1: shift-corrected Tn5 cut-site histogram per strand, and the read counts per region class.
//
// Reference semantics (paths relative to the reference tree):
//   cut site (1-based)   OneRead.get_cutsite                tobias/utils/ngs.pyx:91-106
//   pileup               ReadList.signal + filter           tobias/utils/ngs.pyx:186-199, tools/atacorrect_functions.py:237
//   count_reads          start < cutsite <= end             tools/atacorrect_functions.py:99
//   membership           pysam fetch = alignment overlaps the region (ngs.pyx:152)
//
// One thread per alignment record (grid-stride); the region containing the cut site is found by binary search
// in the sorted region table; counts are integer atomics in L2 (bit-exact, order independent).
#include "tb_common.cuh"

struct tb_reads_view {
    const i64 *lstart;
    const int32_t *rlen;
    const int32_t *clip;
    const uint8_t *flags;
    i64 n;
};

// 0-based linear position of the cut site (= cutsite - 1 in the reference's 1-based convention)
__device__ __forceinline__ i64 tb_cut_pos(const tb_reads_view &r, i64 i, int shift_fwd, int shift_rev, int &rev) {
    uint8_t f = r.flags[i];
    rev = (f & TB_READ_REVERSE) ? 1 : 0;
    i64 s = r.lstart[i];
    if (rev) return s + r.rlen[i] + r.clip[i] + shift_rev;       // (ref_end + clip5 + 1 + shift) - 1
    return s - r.clip[i] + shift_fwd;                            // (ref_start - clip5 + 1 + shift) - 1
}

// mode 0: pileup (start < cut < end) into out[strand][off[j] + pos]; mode 1: count (start < cut <= end);
// mode 2: ReadList.signal (ngs.pyx:186-199): start < cut <= end for EVERY record of the list, whether or not its alignment overlaps the region
template <int MODE>
__global__ void tb_pileup_kernel(tb_reads_view r, int shift_fwd, int shift_rev, i64 n_regions, const i64 *__restrict__ lstart,
                                 const i64 *__restrict__ lend, const i64 *__restrict__ off, u32 *__restrict__ out_fwd,
                                 u32 *__restrict__ out_rev, u64 *__restrict__ counter) {
    i64 stride = (i64)gridDim.x * blockDim.x;
    u32 local = 0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < r.n; i += stride) {
        if (r.flags[i] & TB_READ_SKIP) continue;
        int rev;
        i64 p = tb_cut_pos(r, i, shift_fwd, shift_rev, rev);
        i64 rs = r.lstart[i], re = rs + r.rlen[i];
        // last admissible cut position inside region j: end-2 (pileup) / end-1 (count)
        const i64 slack = (MODE == 0) ? 2 : 1;
        i64 j = tb_upper_bound(lend, n_regions, p + slack - 1);      // first region with lend - slack >= p
        for (; j < n_regions && lstart[j] <= p; j++) {
            if (p > lend[j] - slack) continue;
            if (MODE != 2 && !(rs < lend[j] && re > lstart[j])) continue;         // fetch(): alignment must overlap the region
            if (MODE != 1) {
                u32 *dst = rev ? out_rev : out_fwd;
                atomicAdd(&dst[off[j] + (p - lstart[j])], 1u);
            } else {
                local++;
            }
        }
    }
    if (MODE == 1) {
        for (int d = 16; d > 0; d >>= 1) local += __shfl_down_sync(0xffffffffu, local, d);
        if ((threadIdx.x & 31) == 0 && local) atomicAdd(counter, (u64)local);
    }
}

static int tb_reads_ready(tb_ctx *ctx, tb_reads_view &rv) {
    rv.lstart = ctx->r_start.as<i64>();
    rv.rlen = ctx->r_len.as<int32_t>();
    rv.clip = ctx->r_clip.as<int32_t>();
    rv.flags = ctx->r_flags.as<uint8_t>();
    rv.n = ctx->n_reads;
    return TB_OK;
}

// shared by tb_pileup and tb_correct_run: fills d_out (u32 [2][total]) for regions given in linear coordinates
int tb_pileup_device(tb_ctx *ctx, i64 n_regions, const i64 *d_lstart, const i64 *d_lend, const i64 *d_off, i64 total,
                     int shift_fwd, int shift_rev, u32 *d_out, int list_semantics = 0) {
    tb_reads_view rv;
    tb_reads_ready(ctx, rv);
    TB_CUDA(ctx, cudaMemsetAsync(d_out, 0, (size_t)total * 2 * sizeof(u32), ctx->stream));
    if (rv.n == 0 || n_regions == 0) return TB_OK;
    int block = 256;
    i64 want = tb_div_up(rv.n, block);
    i64 cap = (i64)ctx->sm_count * 16;
    unsigned grid = (unsigned)(want < cap ? want : cap);
    if (list_semantics) {
        auto k = tb_pileup_kernel<2>;
        TB_LAUNCH(ctx, k, grid, block, 0, rv, shift_fwd, shift_rev, n_regions, d_lstart, d_lend, d_off, d_out, d_out + total, (u64 *)nullptr);
    } else {
        auto k = tb_pileup_kernel<0>;
        TB_LAUNCH(ctx, k, grid, block, 0, rv, shift_fwd, shift_rev, n_regions, d_lstart, d_lend, d_off, d_out, d_out + total, (u64 *)nullptr);
    }
    return tb_check(ctx, cudaGetLastError(), "tb_pileup_kernel");
}

static int tb_pileup_any(tb_ctx *ctx, int64_t n_regions, const int32_t *contig, const int64_t *start, const int64_t *end, int shift_fwd,
                         int shift_rev, uint32_t *out_fwd, uint32_t *out_rev, int list_semantics);

extern "C" int tb_pileup(tb_ctx *ctx, int64_t n_regions, const int32_t *contig, const int64_t *start, const int64_t *end,
                         int shift_fwd, int shift_rev, uint32_t *out_fwd, uint32_t *out_rev) {
    return tb_pileup_any(ctx, n_regions, contig, start, end, shift_fwd, shift_rev, out_fwd, out_rev, 0);
}

// ReadList.signal(region) of the resident records (tobias/utils/ngs.pyx:186-199): values[cut - start - 1] += 1 for start < cut <= end
extern "C" int tb_read_signal(tb_ctx *ctx, int64_t n_regions, const int32_t *contig, const int64_t *start, const int64_t *end,
                              int shift_fwd, int shift_rev, uint32_t *out_fwd, uint32_t *out_rev) {
    return tb_pileup_any(ctx, n_regions, contig, start, end, shift_fwd, shift_rev, out_fwd, out_rev, 1);
}

static int tb_pileup_any(tb_ctx *ctx, int64_t n_regions, const int32_t *contig, const int64_t *start, const int64_t *end, int shift_fwd,
                         int shift_rev, uint32_t *out_fwd, uint32_t *out_rev, int list_semantics) {
    if (!ctx) return TB_ERR_ARG;
    std::vector<i64> ls, le;
    TB_TRY(tb_linear_regions(ctx, n_regions, contig, start, end, ls, le, false));
    std::vector<i64> off(n_regions + 1, 0);
    for (i64 i = 0; i < n_regions; i++) off[i + 1] = off[i] + (le[i] - ls[i]);
    i64 total = off[n_regions];
    if (total == 0) return TB_OK;
    TB_TRY(tb_h2d(ctx, ctx->reg_a, ls.data(), sizeof(i64) * n_regions));
    TB_TRY(tb_h2d(ctx, ctx->reg_b, le.data(), sizeof(i64) * n_regions));
    TB_TRY(tb_h2d(ctx, ctx->reg_c, off.data(), sizeof(i64) * (n_regions + 1)));
    TB_TRY(tb_reserve(ctx, ctx->scratch0, (size_t)total * 2 * sizeof(u32)));
    tb_timer_start(ctx);
    TB_TRY(tb_pileup_device(ctx, n_regions, ctx->reg_a.as<i64>(), ctx->reg_b.as<i64>(), ctx->reg_c.as<i64>(), total, shift_fwd,
                            shift_rev, ctx->scratch0.as<u32>(), list_semantics));
    tb_timer_stop(ctx);
    if (out_fwd) TB_TRY(tb_d2h(ctx, out_fwd, ctx->scratch0.as<u32>(), (size_t)total * sizeof(u32)));
    if (out_rev) TB_TRY(tb_d2h(ctx, out_rev, ctx->scratch0.as<u32>() + total, (size_t)total * sizeof(u32)));
    return tb_sync(ctx, "tb_pileup");
}

extern "C" int tb_count_reads(tb_ctx *ctx, int64_t n_regions, const int32_t *contig, const int64_t *start, const int64_t *end,
                              int shift_fwd, int shift_rev, int64_t *count_out) {
    if (!ctx || !count_out) return TB_ERR_ARG;
    std::vector<i64> &ls = ctx->h64[0], &le = ctx->h64[1];
    const tb_tracer tr(ctx, "tb_count_reads");
    TB_TRY(tb_linear_regions(ctx, n_regions, contig, start, end, ls, le, false));
    *count_out = 0;
    tb_reads_view rv;
    tb_reads_ready(ctx, rv);
    if (rv.n == 0 || n_regions == 0) return TB_OK;
    TB_TRY(tb_h2d(ctx, ctx->reg_a, ls.data(), sizeof(i64) * n_regions));
    TB_TRY(tb_h2d(ctx, ctx->reg_b, le.data(), sizeof(i64) * n_regions));
    TB_TRY(tb_reserve(ctx, ctx->reg_d, 64));
    TB_CUDA(ctx, cudaMemsetAsync(ctx->reg_d.p, 0, 64, ctx->stream));
    tr.mark("tables uploaded");
    int block = 256;
    i64 want = tb_div_up(rv.n, block);
    i64 cap = (i64)ctx->sm_count * 16;
    unsigned grid = (unsigned)(want < cap ? want : cap);
    auto k = tb_pileup_kernel<1>;
    tb_timer_start(ctx);
    TB_LAUNCH(ctx, k, grid, block, 0, rv, shift_fwd, shift_rev, (i64)n_regions, ctx->reg_a.as<i64>(), ctx->reg_b.as<i64>(),
              (const i64 *)nullptr, (u32 *)nullptr, (u32 *)nullptr, ctx->reg_d.as<u64>());
    tb_timer_stop(ctx);
    TB_TRY(tb_check(ctx, cudaGetLastError(), "tb_count_kernel"));
    u64 c = 0;
    TB_TRY(tb_d2h(ctx, &c, ctx->reg_d.p, sizeof(u64)));
    *count_out = (int64_t)c;
    return TB_OK;
}
