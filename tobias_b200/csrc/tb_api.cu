// tb_api.cu -- context management, genome packing/upload, read upload, shared host helpers of libtobias_b200.
#include "tb_common.cuh"

static char g_init_error[512] = "";

int tb_fail(tb_ctx *ctx, int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    char *dst = ctx ? ctx->err : g_init_error;
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
}

int tb_check(tb_ctx *ctx, cudaError_t e, const char *what) {
    if (e == cudaSuccess) return TB_OK;
    return tb_fail(ctx, e == cudaErrorMemoryAllocation ? TB_ERR_NOMEM : TB_ERR_CUDA, "CUDA error %d (%s) in %s", (int)e,
                   cudaGetErrorString(e), what);
}

int tb_reserve(tb_ctx *ctx, tb_dbuf &b, size_t bytes) {
    if (bytes <= b.cap && b.p) return TB_OK;
    if (b.p) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(b.p);
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes < 256 ? 256 : bytes;
    want = (want + 255) & ~(size_t)255;
    TB_CUDA(ctx, cudaMalloc(&b.p, want));
    b.cap = want;
    return TB_OK;
}

void tb_release(tb_dbuf &b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

int tb_pinned(tb_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->pinned_cap) return TB_OK;
    if (ctx->pinned) {
        cudaStreamSynchronize(ctx->stream);
        cudaFreeHost(ctx->pinned);
        ctx->pinned = nullptr;
        ctx->pinned_cap = 0;
    }
    TB_CUDA(ctx, cudaMallocHost(&ctx->pinned, bytes));
    ctx->pinned_cap = bytes;
    return TB_OK;
}

int tb_h2d(tb_ctx *ctx, tb_dbuf &dst, const void *src, size_t bytes) {
    TB_TRY(tb_reserve(ctx, dst, bytes));
    if (bytes) TB_CUDA(ctx, cudaMemcpyAsync(dst.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return TB_OK;
}

int tb_h2d_aux(tb_ctx *ctx, tb_dbuf &dst, const void *src, size_t bytes) {
    TB_TRY(tb_reserve(ctx, dst, bytes));
    if (bytes) TB_CUDA(ctx, cudaMemcpyAsync(dst.p, src, bytes, cudaMemcpyHostToDevice, ctx->aux_stream));
    return TB_OK;
}

int tb_aux_join(tb_ctx *ctx) {
    TB_CUDA(ctx, cudaEventRecord(ctx->ev_aux, ctx->aux_stream));
    TB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_aux, 0));
    return TB_OK;
}

int tb_d2h(tb_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (bytes) TB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TB_OK;
}

int tb_sync(tb_ctx *ctx, const char *what) {
    TB_TRY(tb_check(ctx, cudaStreamSynchronize(ctx->stream), what));
    return tb_check(ctx, cudaGetLastError(), what);
}

void tb_timer_start(tb_ctx *ctx) { cudaEventRecord(ctx->ev0, ctx->stream); }
void tb_timer_stop(tb_ctx *ctx) {
    cudaEventRecord(ctx->ev1, ctx->stream);
    cudaEventSynchronize(ctx->ev1);
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
}

void tb_mark(tb_ctx *ctx, int i) { cudaEventRecord(ctx->marks[i], ctx->stream); }
void tb_marks_collect(tb_ctx *ctx, int n_marks) {
    cudaEventSynchronize(ctx->marks[n_marks - 1]);
    for (int i = 0; i < 8; i++) ctx->stage_ms[i] = 0.f;
    for (int i = 0; i + 1 < n_marks; i++) cudaEventElapsedTime(&ctx->stage_ms[i], ctx->marks[i], ctx->marks[i + 1]);
}

int tb_linear_regions(tb_ctx *ctx, i64 n, const int32_t *contig, const int64_t *start, const int64_t *end,
                      std::vector<i64> &lstart, std::vector<i64> &lend, bool need_disjoint) {
    if (ctx->n_contigs == 0) return tb_fail(ctx, TB_ERR_STATE, "no genome: call tb_genome_create first");
    if (n < 0) return tb_fail(ctx, TB_ERR_ARG, "negative region count");
    lstart.resize(n);
    lend.resize(n);
    for (i64 i = 0; i < n; i++) {
        int c = contig[i];
        if (c < 0 || c >= ctx->n_contigs) return tb_fail(ctx, TB_ERR_ARG, "region %lld: contig %d out of range", i, c);
        if (start[i] < 0 || end[i] > ctx->contig_len[c] || start[i] >= end[i])
            return tb_fail(ctx, TB_ERR_ARG, "region %lld: [%lld, %lld) is empty or outside contig %d (length %lld)", i,
                           (i64)start[i], (i64)end[i], c, ctx->contig_len[c]);
        lstart[i] = ctx->contig_off[c] + start[i];
        lend[i] = ctx->contig_off[c] + end[i];
        if (i > 0) {
            if (lstart[i] < lstart[i - 1] || lend[i] < lend[i - 1])
                return tb_fail(ctx, TB_ERR_ARG, "regions must be sorted by (contig, start) with ascending ends (region %lld)", i);
            if (need_disjoint && lstart[i] < lend[i - 1])
                return tb_fail(ctx, TB_ERR_ARG, "regions must not overlap (region %lld)", i);
        }
    }
    return TB_OK;
}

// ------------------------------------------------------------------------------------------------ context
extern "C" int tb_abi_version(void) { return 1; }

extern "C" int tb_init(int device, tb_ctx **out) {
    if (!out) return tb_fail(nullptr, TB_ERR_ARG, "tb_init: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return tb_fail(nullptr, TB_ERR_CUDA, "tb_init: no usable CUDA device (%s); libtobias_b200 has no CPU fallback",
                       e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return tb_fail(nullptr, TB_ERR_ARG, "tb_init: device %d out of range (%d devices)", device, ndev);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return tb_fail(nullptr, TB_ERR_CUDA, "tb_init: cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    tb_ctx *ctx = new tb_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        delete ctx;
        return tb_fail(nullptr, TB_ERR_CUDA, "tb_init: cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    }
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithPriority(&ctx->copy_stream, cudaStreamNonBlocking, -1) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_aux, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming) != cudaSuccess || cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) {
        delete ctx;
        return tb_fail(nullptr, TB_ERR_CUDA, "tb_init: stream/event creation failed");
    }
    bool marks_ok = true;
    for (int i = 0; i < 8; i++) marks_ok = marks_ok && cudaEventCreate(&ctx->marks[i]) == cudaSuccess;
    for (int i = 0; i < 8; i++) marks_ok = marks_ok && cudaEventCreateWithFlags(&ctx->ev_chunk[i], cudaEventDisableTiming) == cudaSuccess;
    if (!marks_ok) {
        delete ctx;
        return tb_fail(nullptr, TB_ERR_CUDA, "tb_init: stream/event creation failed");
    }
    *out = ctx;
    return TB_OK;
}

// Switches between kernel variants / chunk sizes that produce the same results (comparison runs and tests; never needed in production).
extern "C" int tb_set_option(tb_ctx *ctx, const char *key, int64_t value) {
    if (!ctx || !key) return TB_ERR_ARG;
    static const char *known[] = {"k3a_sum_form", "k3a_no_tmem", "k3a_tmem_groups", "k3c_fast", "k3c_chunks", "k4_chunks", "k2_tile", "k2_site_cap", "k2_chunk",
                                  "k2_no_hist", "k2_smem_chunk", "k2_no_smem_hist", "k4_form", "k4_threads", "k3b_persistent", "k3b_tile", "k3b_cap", "k3b_threads", "host_trace", nullptr};
    bool ok = false;
    for (int i = 0; known[i]; i++) ok = ok || strcmp(known[i], key) == 0;
    if (!ok) return tb_fail(ctx, TB_ERR_ARG, "tb_set_option: unknown option '%s'", key);
    if (value < 0) ctx->opts.erase(key);          // negative: back to the default
    else ctx->opts[key] = value;
    ctx->force_sum_form = (int)tb_opt(ctx, "k3a_sum_form", 0);
    ctx->k3a_no_tmem = (int)tb_opt(ctx, "k3a_no_tmem", 0);
    ctx->k3a_tmem_groups = (int)tb_opt(ctx, "k3a_tmem_groups", 0);
    return TB_OK;
}

extern "C" void tb_destroy(tb_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copy_stream);
    for (int t = 0; t < 4; t++)
        for (int s2 = 0; s2 < 2; s2++) tb_release(ctx->cor.dl[t][s2]);
    tb_dbuf *bufs[] = {&ctx->seq2, &ctx->nmask, &ctx->d_contig_off, &ctx->r_start, &ctx->r_len, &ctx->r_clip, &ctx->r_flags,
                       &ctx->reg_a, &ctx->reg_b, &ctx->reg_c, &ctx->reg_d, &ctx->reg_e, &ctx->scratch0, &ctx->scratch1,
                       &ctx->scratch2, &ctx->scratch3, &ctx->scratch4, &ctx->staging, &ctx->model[0].table, &ctx->model[1].table,
                       &ctx->model[0].vtable, &ctx->model[1].vtable, &ctx->k2_sites, &ctx->k2_hist, &ctx->k2_wts, &ctx->staging_up,
                       &ctx->cor.d_ext_start, &ctx->cor.d_ext_end, &ctx->cor.d_ext_off, &ctx->cor.d_out_off, &ctx->cor.d_win_off,
                       &ctx->cor.pile, &ctx->cor.biasraw, &ctx->cor.fit, &ctx->cor.fitg_x, &ctx->cor.fitg_y, &ctx->cor.fitg_st,
                       &ctx->cor.fitg_sti, &ctx->cor.fitg_wq, &ctx->cor.fitg_lists, &ctx->cor.fitg_sw, &ctx->cor.tracks, &ctx->cor.prepost,
                       &ctx->sc.d_ext_off, &ctx->sc.signal, &ctx->sc.work0, &ctx->sc.work1, &ctx->sc.out};
    for (tb_dbuf *b : bufs) tb_release(*b);
    for (tb_dbuf &b : ctx->nrun_buf) tb_release(b);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->cor.pp_pinned) cudaFreeHost(ctx->cor.pp_pinned);
    tb_nccl_teardown(ctx);
    for (int i = 0; i < 8; i++) cudaEventDestroy(ctx->marks[i]);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    cudaEventDestroy(ctx->ev_copy);
    if (ctx->ev_aux) cudaEventDestroy(ctx->ev_aux);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    for (auto &e : ctx->ev_chunk)
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->contig_ready)
        if (e) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->copy_stream);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char *tb_last_error(const tb_ctx *ctx) { return ctx ? ctx->err : g_init_error; }

extern "C" int tb_device_info(tb_ctx *ctx, int *sm_count, int64_t *free_bytes, int64_t *total_bytes, int *is_emulator) {
    if (!ctx) return TB_ERR_ARG;
    size_t f = 0, t = 0;
    TB_CUDA(ctx, cudaMemGetInfo(&f, &t));
    if (sm_count) *sm_count = ctx->sm_count;
    if (free_bytes) *free_bytes = (int64_t)f;
    if (total_bytes) *total_bytes = (int64_t)t;
#ifdef TB_EMU
    if (is_emulator) *is_emulator = 1;
#else
    if (is_emulator) *is_emulator = 0;
#endif
    return TB_OK;
}

extern "C" int tb_last_kernel_ms(tb_ctx *ctx, float *ms) {
    if (!ctx || !ms) return TB_ERR_ARG;
    *ms = ctx->last_ms;
    return TB_OK;
}

extern "C" int tb_timer_begin(tb_ctx *ctx) {
    if (!ctx) return TB_ERR_ARG;
    return tb_check(ctx, cudaEventRecord(ctx->marks[6], ctx->stream), "tb_timer_begin");
}

extern "C" int tb_timer_end(tb_ctx *ctx, float *ms) {
    if (!ctx || !ms) return TB_ERR_ARG;
    TB_CUDA(ctx, cudaEventRecord(ctx->marks[7], ctx->stream));
    TB_CUDA(ctx, cudaEventSynchronize(ctx->marks[7]));
    return tb_check(ctx, cudaEventElapsedTime(ms, ctx->marks[6], ctx->marks[7]), "tb_timer_end");
}

extern "C" int tb_host_alloc(tb_ctx *ctx, int64_t bytes, void **out) {
    if (!ctx || !out || bytes < 0) return TB_ERR_ARG;
    return tb_check(ctx, cudaMallocHost(out, (size_t)(bytes > 0 ? bytes : 1)), "tb_host_alloc");
}

extern "C" int tb_host_free(tb_ctx *ctx, void *p) {
    if (!ctx) return TB_ERR_ARG;
    return p ? tb_check(ctx, cudaFreeHost(p), "tb_host_free") : TB_OK;
}

extern "C" int tb_stage_ms(tb_ctx *ctx, float *out, int n) {
    if (!ctx || !out || n < 0) return TB_ERR_ARG;
    for (int i = 0; i < n; i++) out[i] = i < 8 ? ctx->stage_ms[i] : 0.f;
    return TB_OK;
}

extern "C" int tb_launch_count(tb_ctx *ctx, int64_t *n) {
    if (!ctx || !n) return TB_ERR_ARG;
    *n = ctx->launches;
    return TB_OK;
}

// ------------------------------------------------------------------------------------------------ genome
// ASCII -> (2-bit code, N flag).  One thread packs 32 bases: two u32 words of seq2 and one u32 word of nmask.
// 128-bit loads of the ASCII staging buffer (32 bases = 2 x uint4), coalesced 4/8-byte stores.
__device__ __forceinline__ u32 tb_ascii_code(u32 c, u32 &is_n) {
    c &= 0xDFu;                       // upper-case
    u32 code = 0;
    is_n = 0;
    if (c == 'A') code = 0;
    else if (c == 'T') code = 1;
    else if (c == 'C') code = 2;
    else if (c == 'G') code = 3;
    else is_n = 1;
    return code;
}

__global__ void tb_pack_kernel(const uint8_t *__restrict__ ascii, i64 n, u32 *__restrict__ seq2, u32 *__restrict__ nmask,
                               i64 base0) {
    i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;     // group of 32 bases
    i64 first = g * 32;
    if (first >= n) return;
    u32 w0 = 0, w1 = 0, m = 0;
    if (first + 32 <= n) {
        const uint4 *p = (const uint4 *)(ascii + first);
        uint4 a = p[0], b = p[1];
        u32 words[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 32; i++) {
            u32 c = (words[i >> 2] >> (8 * (i & 3))) & 0xFFu, isn;
            u32 code = tb_ascii_code(c, isn);
            if (i < 16) w0 |= code << (2 * i); else w1 |= code << (2 * (i - 16));
            m |= isn << i;
        }
    } else {
        for (int i = 0; i < 32; i++) {
            u32 isn = 1, code = 0;
            if (first + i < n) code = tb_ascii_code(ascii[first + i], isn);
            if (i < 16) w0 |= code << (2 * i); else w1 |= code << (2 * (i - 16));
            m |= isn << i;
        }
    }
    i64 k = base0 + first;            // multiple of 32
    seq2[(k >> 4)] = w0;
    seq2[(k >> 4) + 1] = w1;
    nmask[k >> 5] = m;
}

__global__ void tb_fetch_codes_kernel(tb_genome_view g, i64 k0, i64 n, uint8_t *out) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = tb_is_n(g, k0 + i) ? 4 : (uint8_t)tb_base(g, k0 + i);
}

static const i64 TB_CONTIG_PAD = 1024;     // >= this many N bases between contigs in the linear space

extern "C" int tb_genome_create(tb_ctx *ctx, int n_contigs, const int64_t *lengths) {
    if (!ctx) return TB_ERR_ARG;
    if (n_contigs <= 0 || !lengths) return tb_fail(ctx, TB_ERR_ARG, "tb_genome_create: need at least one contig");
    TB_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));           // asynchronous uploads into the previous genome
    ctx->contig_pending.assign(n_contigs, 0);
    ctx->contig_len.assign(lengths, lengths + n_contigs);
    ctx->contig_off.resize(n_contigs + 1);
    i64 off = TB_CONTIG_PAD;
    for (int c = 0; c < n_contigs; c++) {
        if (lengths[c] <= 0) return tb_fail(ctx, TB_ERR_ARG, "tb_genome_create: contig %d has length %lld", c, (i64)lengths[c]);
        ctx->contig_off[c] = off;
        off += lengths[c] + TB_CONTIG_PAD;
        off = (off + 255) & ~(i64)255;
    }
    ctx->contig_off[n_contigs] = off;
    ctx->genome_bases = off;
    ctx->n_contigs = n_contigs;
    size_t seq_bytes = (size_t)(off / 16 + 8) * 4, mask_bytes = (size_t)(off / 32 + 8) * 4;
    TB_TRY(tb_reserve(ctx, ctx->seq2, seq_bytes));
    TB_TRY(tb_reserve(ctx, ctx->nmask, mask_bytes));
    TB_CUDA(ctx, cudaMemsetAsync(ctx->seq2.p, 0, seq_bytes, ctx->stream));
    TB_CUDA(ctx, cudaMemsetAsync(ctx->nmask.p, 0xFF, mask_bytes, ctx->stream));     // everything is N until uploaded
    TB_TRY(tb_h2d(ctx, ctx->d_contig_off, ctx->contig_off.data(), sizeof(i64) * (n_contigs + 1)));
    ctx->cor.valid = false;
    return tb_sync(ctx, "tb_genome_create");
}

extern "C" int tb_genome_upload(tb_ctx *ctx, int contig, int64_t offset, const uint8_t *ascii, int64_t n) {
    if (!ctx) return TB_ERR_ARG;
    if (ctx->n_contigs == 0) return tb_fail(ctx, TB_ERR_STATE, "tb_genome_upload: call tb_genome_create first");
    if (contig < 0 || contig >= ctx->n_contigs) return tb_fail(ctx, TB_ERR_ARG, "tb_genome_upload: contig %d out of range", contig);
    if (offset < 0 || (offset & 31) || n < 0 || offset + n > ctx->contig_len[contig])
        return tb_fail(ctx, TB_ERR_ARG, "tb_genome_upload: bad range offset=%lld n=%lld (offset must be a multiple of 32)",
                       (i64)offset, (i64)n);
    TB_TRY(tb_genome_fence(ctx, contig, contig));                    // after an asynchronous upload of the same contig
    const i64 CHUNK = (i64)64 << 20;
    for (i64 done = 0; done < n; done += CHUNK) {
        i64 m = n - done < CHUNK ? n - done : CHUNK;
        TB_TRY(tb_reserve(ctx, ctx->staging, (size_t)m + 64));
        TB_CUDA(ctx, cudaMemcpyAsync(ctx->staging.p, ascii + done, (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
        i64 groups = tb_div_up(m, 32);
        int block = 256;
        TB_LAUNCH(ctx, tb_pack_kernel, (unsigned)tb_div_up(groups, block), block, 0, ctx->staging.as<uint8_t>(), m,
                  ctx->seq2.as<u32>(), ctx->nmask.as<u32>(), ctx->contig_off[contig] + offset + done);
        TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));     // the host buffer may be reused by the caller
    }
    return tb_check(ctx, cudaGetLastError(), "tb_pack_kernel");
}

// Asynchronous form: the copies and the packing kernels are enqueued on the copy stream and the call returns; `ascii` (page-locked,
// tb_host_alloc, for a truly asynchronous copy) must stay untouched until tb_genome_wait.  Everything that reads the genome waits on
// the device for the contigs it needs, so e.g. tb_bias_counts works on the first contigs while the later ones are still in flight.
extern "C" int tb_genome_upload_async(tb_ctx *ctx, int contig, int64_t offset, const uint8_t *ascii, int64_t n) {
    if (!ctx) return TB_ERR_ARG;
    if (ctx->n_contigs == 0) return tb_fail(ctx, TB_ERR_STATE, "tb_genome_upload_async: call tb_genome_create first");
    if (contig < 0 || contig >= ctx->n_contigs) return tb_fail(ctx, TB_ERR_ARG, "tb_genome_upload_async: contig %d out of range", contig);
    if (offset < 0 || (offset & 31) || n < 0 || offset + n > ctx->contig_len[contig])
        return tb_fail(ctx, TB_ERR_ARG, "tb_genome_upload_async: bad range offset=%lld n=%lld (offset must be a multiple of 32)",
                       (i64)offset, (i64)n);
    const i64 CHUNK = (i64)64 << 20;
    if (ctx->staging_up.cap < (size_t)CHUNK + 64) {
        TB_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));       // nothing in flight may still read the old buffer
        TB_TRY(tb_reserve(ctx, ctx->staging_up, (size_t)CHUNK + 64));
    }
    if ((int)ctx->contig_ready.size() < ctx->n_contigs) {
        size_t have = ctx->contig_ready.size();
        ctx->contig_ready.resize(ctx->n_contigs, nullptr);
        for (size_t c = have; c < ctx->contig_ready.size(); c++)
            TB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->contig_ready[c], cudaEventDisableTiming));
    }
    ctx->contig_pending.resize(ctx->n_contigs, 0);
    for (i64 done = 0; done < n; done += CHUNK) {                    // one stream, one staging buffer: in order, no hazards
        i64 m = n - done < CHUNK ? n - done : CHUNK;
        TB_CUDA(ctx, cudaMemcpyAsync(ctx->staging_up.p, ascii + done, (size_t)m, cudaMemcpyHostToDevice, ctx->copy_stream));
        TB_LAUNCH_ON(ctx, ctx->copy_stream, tb_pack_kernel, (unsigned)tb_div_up(tb_div_up(m, 32), 256), 256, 0, ctx->staging_up.as<uint8_t>(), m,
                     ctx->seq2.as<u32>(), ctx->nmask.as<u32>(), ctx->contig_off[contig] + offset + done);
    }
    TB_CUDA(ctx, cudaEventRecord(ctx->contig_ready[contig], ctx->copy_stream));
    ctx->contig_pending[contig] = 1;
    return tb_check(ctx, cudaGetLastError(), "tb_genome_upload_async");
}

// Host-side completion of every asynchronous upload (the host buffers may be reused afterwards).
// A contig already in the device layout (host packer: tobias_b200/io/hostpack.c): two plain copies straight into place -- contig offsets
// are multiples of 256 bases, so word boundaries agree.  async != 0: on the copy stream, like tb_genome_upload_async.
extern "C" int tb_genome_upload_packed(tb_ctx *ctx, int contig, const uint32_t *seq2, const uint32_t *nmask, int64_t n_bases, int async) {
    if (!ctx || !seq2 || !nmask) return TB_ERR_ARG;
    if (ctx->n_contigs == 0) return tb_fail(ctx, TB_ERR_STATE, "tb_genome_upload_packed: call tb_genome_create first");
    if (contig < 0 || contig >= ctx->n_contigs) return tb_fail(ctx, TB_ERR_ARG, "tb_genome_upload_packed: contig %d out of range", contig);
    if (n_bases != ctx->contig_len[contig])
        return tb_fail(ctx, TB_ERR_ARG, "tb_genome_upload_packed: contig %d has %lld bases, got %lld", contig, (i64)ctx->contig_len[contig], (i64)n_bases);
    const i64 base = ctx->contig_off[contig];
    u32 *d_seq = ctx->seq2.as<u32>() + (base >> 4), *d_mask = ctx->nmask.as<u32>() + (base >> 5);
    const size_t b_seq = (size_t)((n_bases + 15) / 16) * sizeof(u32), b_mask = (size_t)((n_bases + 31) / 32) * sizeof(u32);
    if (!async) {
        TB_TRY(tb_genome_fence(ctx, contig, contig));
        TB_CUDA(ctx, cudaMemcpyAsync(d_seq, seq2, b_seq, cudaMemcpyHostToDevice, ctx->stream));
        TB_CUDA(ctx, cudaMemcpyAsync(d_mask, nmask, b_mask, cudaMemcpyHostToDevice, ctx->stream));
        return tb_sync(ctx, "tb_genome_upload_packed");
    }
    if ((int)ctx->contig_ready.size() < ctx->n_contigs) {
        size_t have = ctx->contig_ready.size();
        ctx->contig_ready.resize(ctx->n_contigs, nullptr);
        for (size_t c = have; c < ctx->contig_ready.size(); c++)
            TB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->contig_ready[c], cudaEventDisableTiming));
    }
    ctx->contig_pending.resize(ctx->n_contigs, 0);
    TB_CUDA(ctx, cudaMemcpyAsync(d_seq, seq2, b_seq, cudaMemcpyHostToDevice, ctx->copy_stream));
    TB_CUDA(ctx, cudaMemcpyAsync(d_mask, nmask, b_mask, cudaMemcpyHostToDevice, ctx->copy_stream));
    TB_CUDA(ctx, cudaEventRecord(ctx->contig_ready[contig], ctx->copy_stream));
    ctx->contig_pending[contig] = 1;
    return tb_check(ctx, cudaGetLastError(), "tb_genome_upload_packed");
}

// bits [s, e) of a bit mask (32 positions per word) set by one warp per run; runs may share a word
__global__ void tb_mask_runs_kernel(u32 *__restrict__ mask, const i64 *__restrict__ run_start, const i64 *__restrict__ run_end, i64 n_runs) {
    const i64 r = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n_runs) return;
    const i64 s = run_start[r], e = run_end[r];
    for (i64 w = (s >> 5) + lane; w <= ((e - 1) >> 5); w += 32) {
        const i64 lo = w << 5;
        const int b0 = s > lo ? (int)(s - lo) : 0, b1 = e < lo + 32 ? (int)(e - lo) : 32;
        const u32 bits = (b1 - b0 == 32) ? 0xffffffffu : (((1u << (b1 - b0)) - 1u) << b0);
        atomicOr(&mask[w], bits);
    }
}

extern "C" int tb_genome_upload_packed_runs(tb_ctx *ctx, int contig, const uint32_t *seq2, int64_t n_runs, const int64_t *run_start,
                                            const int64_t *run_end, int64_t n_bases, int async) {
    if (!ctx || !seq2 || n_runs < 0 || (n_runs > 0 && (!run_start || !run_end))) return TB_ERR_ARG;
    if (ctx->n_contigs == 0) return tb_fail(ctx, TB_ERR_STATE, "tb_genome_upload_packed_runs: call tb_genome_create first");
    if (contig < 0 || contig >= ctx->n_contigs) return tb_fail(ctx, TB_ERR_ARG, "tb_genome_upload_packed_runs: contig %d out of range", contig);
    if (n_bases != ctx->contig_len[contig])
        return tb_fail(ctx, TB_ERR_ARG, "tb_genome_upload_packed_runs: contig %d has %lld bases, got %lld", contig, (i64)ctx->contig_len[contig], (i64)n_bases);
    for (i64 r = 0; r < n_runs; r++)
        if (run_start[r] < 0 || run_end[r] <= run_start[r] || run_end[r] > n_bases)
            return tb_fail(ctx, TB_ERR_ARG, "tb_genome_upload_packed_runs: run %lld [%lld, %lld) is empty or outside the contig", r, (i64)run_start[r], (i64)run_end[r]);
    const i64 base = ctx->contig_off[contig];
    u32 *d_seq = ctx->seq2.as<u32>() + (base >> 4), *d_mask = ctx->nmask.as<u32>() + (base >> 5);
    const size_t b_seq = (size_t)((n_bases + 15) / 16) * sizeof(u32), b_mask = (size_t)((n_bases + 31) / 32) * sizeof(u32);
    cudaStream_t st = async ? ctx->copy_stream : ctx->stream;
    if (async) {
        if ((int)ctx->contig_ready.size() < ctx->n_contigs) {
            size_t have = ctx->contig_ready.size();
            ctx->contig_ready.resize(ctx->n_contigs, nullptr);
            for (size_t c = have; c < ctx->contig_ready.size(); c++)
                TB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->contig_ready[c], cudaEventDisableTiming));
        }
        ctx->contig_pending.resize(ctx->n_contigs, 0);
    } else {
        TB_TRY(tb_genome_fence(ctx, contig, contig));
    }
    TB_CUDA(ctx, cudaMemcpyAsync(d_seq, seq2, b_seq, cudaMemcpyHostToDevice, st));
    TB_CUDA(ctx, cudaMemsetAsync(d_mask, 0, b_mask, st));
    if (n_runs > 0) {
        // the run table travels through a buffer of its own per contig (the copies of several contigs are in flight at once)
        if ((int)ctx->nrun_buf.size() < ctx->n_contigs) ctx->nrun_buf.resize(ctx->n_contigs);
        tb_dbuf &rb = ctx->nrun_buf[contig];
        TB_TRY(tb_reserve(ctx, rb, (size_t)2 * n_runs * sizeof(i64)));
        TB_CUDA(ctx, cudaMemcpyAsync(rb.p, run_start, (size_t)n_runs * sizeof(i64), cudaMemcpyHostToDevice, st));
        TB_CUDA(ctx, cudaMemcpyAsync(rb.as<i64>() + n_runs, run_end, (size_t)n_runs * sizeof(i64), cudaMemcpyHostToDevice, st));
        TB_LAUNCH_ON(ctx, st, tb_mask_runs_kernel, (unsigned)tb_div_up(n_runs * 32, 256), 256, 0, d_mask, (const i64 *)rb.as<i64>(),
                     (const i64 *)(rb.as<i64>() + n_runs), (i64)n_runs);
    }
    if (async) {
        TB_CUDA(ctx, cudaEventRecord(ctx->contig_ready[contig], ctx->copy_stream));
        ctx->contig_pending[contig] = 1;
        return tb_check(ctx, cudaGetLastError(), "tb_genome_upload_packed_runs");
    }
    return tb_sync(ctx, "tb_genome_upload_packed_runs");
}

extern "C" int tb_genome_wait(tb_ctx *ctx) {
    if (!ctx) return TB_ERR_ARG;
    TB_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    ctx->contig_pending.assign(ctx->contig_pending.size(), 0);
    return tb_check(ctx, cudaGetLastError(), "tb_genome_wait");
}

int tb_genome_fence(tb_ctx *ctx, int c_lo, int c_hi) {
    const int n = (int)ctx->contig_pending.size();
    for (int c = c_lo < 0 ? 0 : c_lo; c <= c_hi && c < n; c++)
        if (ctx->contig_pending[c]) TB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->contig_ready[c], 0));
    return TB_OK;
}

extern "C" int tb_genome_fetch(tb_ctx *ctx, int contig, int64_t start, int64_t n, uint8_t *codes_out) {
    if (!ctx) return TB_ERR_ARG;
    if (contig < 0 || contig >= ctx->n_contigs || start < 0 || n < 0 || start + n > ctx->contig_len[contig])
        return tb_fail(ctx, TB_ERR_ARG, "tb_genome_fetch: bad range");
    if (n == 0) return TB_OK;
    TB_TRY(tb_genome_fence(ctx, contig, contig));
    TB_TRY(tb_reserve(ctx, ctx->scratch0, (size_t)n));
    tb_genome_view g{ctx->seq2.as<u32>(), ctx->nmask.as<u32>(), ctx->genome_bases};
    TB_LAUNCH(ctx, tb_fetch_codes_kernel, (unsigned)tb_div_up(n, 256), 256, 0, g, ctx->contig_off[contig] + start, (i64)n,
              ctx->scratch0.as<uint8_t>());
    return tb_d2h(ctx, codes_out, ctx->scratch0.p, (size_t)n);
}

// ------------------------------------------------------------------------------------------------ reads
// Host columns are int32 contig-local; the device keeps the linear start (int64), the reference length and
// the 5' clip so that every kernel works in one coordinate space.
__global__ void tb_reads_linearise_kernel(i64 n, const int32_t *__restrict__ contig, const int32_t *__restrict__ ref_start,
                                          const int32_t *__restrict__ ref_end, const i64 *__restrict__ contig_off,
                                          int n_contigs, i64 *__restrict__ lstart, int32_t *__restrict__ rlen,
                                          uint8_t *__restrict__ flags, int *__restrict__ stats) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = contig[i];
    // stats[0] = longest reference span, stats[1] = 1 when the records are not sorted by (contig, start)
    if (i > 0 && (contig[i - 1] > c || (contig[i - 1] == c && ref_start[i - 1] > ref_start[i]))) stats[1] = 1;
    if (c < 0 || c >= n_contigs) {      // unknown contig: never matches any region
        lstart[i] = -((i64)1 << 40);
        rlen[i] = 0;
        flags[i] = (uint8_t)(flags[i] | TB_READ_SKIP);
        return;
    }
    lstart[i] = contig_off[c] + (i64)ref_start[i];
    rlen[i] = ref_end[i] - ref_start[i];
    if (rlen[i] > stats[0]) atomicMax(&stats[0], rlen[i]);
}

// Compact columns: the records of contig c are [rec_off[c], rec_off[c + 1]); start int32, reference length and 5' clip uint16.
__global__ void tb_reads_expand_kernel(i64 n, const i64 *__restrict__ rec_off, int n_contigs, const int32_t *__restrict__ ref_start,
                                       const unsigned short *__restrict__ len16, const unsigned short *__restrict__ clip16,
                                       const i64 *__restrict__ contig_off, i64 *__restrict__ lstart, int32_t *__restrict__ rlen,
                                       int32_t *__restrict__ rclip, int *__restrict__ stats) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = (int)(tb_upper_bound(rec_off, (i64)n_contigs + 1, i) - 1);
    if (i > rec_off[c] && ref_start[i - 1] > ref_start[i]) stats[1] = 1;
    lstart[i] = contig_off[c] + (i64)ref_start[i];
    const int l = (int)len16[i];
    rlen[i] = l;
    rclip[i] = (int)clip16[i];
    if (l > stats[0]) atomicMax(&stats[0], l);
}

extern "C" int tb_reads_upload_compact(tb_ctx *ctx, int64_t n, const int64_t *rec_off, const int32_t *ref_start, const uint16_t *ref_len,
                                       const uint16_t *clip5, const uint8_t *flags) {
    if (!ctx) return TB_ERR_ARG;
    if (ctx->n_contigs == 0) return tb_fail(ctx, TB_ERR_STATE, "tb_reads_upload_compact: call tb_genome_create first");
    if (n < 0 || !rec_off || (n > 0 && (!ref_start || !ref_len || !clip5 || !flags)))
        return tb_fail(ctx, TB_ERR_ARG, "tb_reads_upload_compact: NULL column");
    if (rec_off[0] != 0 || rec_off[ctx->n_contigs] != n) return tb_fail(ctx, TB_ERR_ARG, "tb_reads_upload_compact: rec_off must run from 0 to n");
    for (int c = 0; c < ctx->n_contigs; c++)
        if (rec_off[c + 1] < rec_off[c]) return tb_fail(ctx, TB_ERR_ARG, "tb_reads_upload_compact: rec_off must not decrease (contig %d)", c);
    ctx->n_reads = n;
    ctx->cor.valid = false;
    if (n == 0) return TB_OK;
    TB_TRY(tb_h2d(ctx, ctx->scratch0, rec_off, sizeof(i64) * ((size_t)ctx->n_contigs + 1)));
    TB_TRY(tb_h2d(ctx, ctx->scratch1, ref_start, (size_t)n * 4));
    TB_TRY(tb_h2d(ctx, ctx->scratch2, ref_len, (size_t)n * 2));
    TB_TRY(tb_h2d(ctx, ctx->scratch3, clip5, (size_t)n * 2));
    TB_TRY(tb_h2d(ctx, ctx->r_flags, flags, (size_t)n));
    TB_TRY(tb_reserve(ctx, ctx->r_start, (size_t)n * 8));
    TB_TRY(tb_reserve(ctx, ctx->r_len, (size_t)n * 4));
    TB_TRY(tb_reserve(ctx, ctx->r_clip, (size_t)n * 4));
    TB_TRY(tb_reserve(ctx, ctx->reg_e, 2 * sizeof(int)));
    TB_CUDA(ctx, cudaMemsetAsync(ctx->reg_e.p, 0, 2 * sizeof(int), ctx->stream));
    TB_LAUNCH(ctx, tb_reads_expand_kernel, (unsigned)tb_div_up(n, 256), 256, 0, (i64)n, (const i64 *)ctx->scratch0.as<i64>(), ctx->n_contigs,
              (const int32_t *)ctx->scratch1.as<int32_t>(), (const unsigned short *)ctx->scratch2.as<unsigned short>(),
              (const unsigned short *)ctx->scratch3.as<unsigned short>(), (const i64 *)ctx->d_contig_off.as<i64>(), ctx->r_start.as<i64>(),
              ctx->r_len.as<int32_t>(), ctx->r_clip.as<int32_t>(), ctx->reg_e.as<int>());
    int stats[2] = {0, 0};
    TB_TRY(tb_d2h(ctx, stats, ctx->reg_e.p, sizeof(stats)));
    ctx->max_read_len = stats[0];
    ctx->reads_sorted = stats[1] == 0;
    return tb_sync(ctx, "tb_reads_upload_compact");
}

extern "C" int tb_reads_upload(tb_ctx *ctx, int64_t n, const int32_t *contig, const int32_t *ref_start,
                               const int32_t *ref_end, const int32_t *clip5, const uint8_t *flags) {
    if (!ctx) return TB_ERR_ARG;
    if (ctx->n_contigs == 0) return tb_fail(ctx, TB_ERR_STATE, "tb_reads_upload: call tb_genome_create first");
    if (n < 0 || (n > 0 && (!contig || !ref_start || !ref_end || !clip5 || !flags)))
        return tb_fail(ctx, TB_ERR_ARG, "tb_reads_upload: NULL column");
    ctx->n_reads = n;
    ctx->cor.valid = false;
    if (n == 0) return TB_OK;
    size_t b4 = (size_t)n * 4;
    TB_TRY(tb_h2d(ctx, ctx->scratch0, contig, b4));
    TB_TRY(tb_h2d(ctx, ctx->scratch1, ref_start, b4));
    TB_TRY(tb_h2d(ctx, ctx->scratch2, ref_end, b4));
    TB_TRY(tb_h2d(ctx, ctx->r_clip, clip5, b4));
    TB_TRY(tb_h2d(ctx, ctx->r_flags, flags, (size_t)n));
    TB_TRY(tb_reserve(ctx, ctx->r_start, (size_t)n * 8));
    TB_TRY(tb_reserve(ctx, ctx->r_len, b4));
    TB_TRY(tb_reserve(ctx, ctx->reg_e, 2 * sizeof(int)));
    TB_CUDA(ctx, cudaMemsetAsync(ctx->reg_e.p, 0, 2 * sizeof(int), ctx->stream));
    TB_LAUNCH(ctx, tb_reads_linearise_kernel, (unsigned)tb_div_up(n, 256), 256, 0, (i64)n, ctx->scratch0.as<int32_t>(),
              ctx->scratch1.as<int32_t>(), ctx->scratch2.as<int32_t>(), ctx->d_contig_off.as<i64>(), ctx->n_contigs,
              ctx->r_start.as<i64>(), ctx->r_len.as<int32_t>(), ctx->r_flags.as<uint8_t>(), ctx->reg_e.as<int>());
    int stats[2] = {0, 0};
    TB_TRY(tb_d2h(ctx, stats, ctx->reg_e.p, sizeof(stats)));
    ctx->max_read_len = stats[0];
    ctx->reads_sorted = stats[1] == 0;
    return tb_sync(ctx, "tb_reads_upload");
}
