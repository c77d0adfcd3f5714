// tb_common.cuh -- shared declarations of libtobias_b200 (context, device buffers, launch macro, genome access).
#pragma once

#ifdef TB_EMU
#include "cuda_emu.h"      // tests/emu: GPU-less kernel-logic emulator used by the CPU test-suite only
#define TB_LAUNCH(ctx, kernel, grid, block, smem, ...)                                   \
    do {                                                                                 \
        (ctx)->launches++;                                                               \
        tb_emu::launch(dim3(grid), dim3(block), (size_t)(smem), [&] { kernel(__VA_ARGS__); }); \
    } while (0)
#define TB_LAUNCH_ON(ctx, strm, kernel, grid, block, smem, ...) TB_LAUNCH(ctx, kernel, grid, block, smem, __VA_ARGS__)
#define TB_DYN_SMEM(name) unsigned char *name = tb_emu::S().dyn_smem
#else
#include <cuda_runtime.h>
#define TB_LAUNCH(ctx, kernel, grid, block, smem, ...)                                   \
    do {                                                                                 \
        (ctx)->launches++;                                                               \
        kernel<<<dim3(grid), dim3(block), (size_t)(smem), (ctx)->stream>>>(__VA_ARGS__); \
    } while (0)
#define TB_LAUNCH_ON(ctx, strm, kernel, grid, block, smem, ...)                          \
    do {                                                                                 \
        (ctx)->launches++;                                                               \
        kernel<<<dim3(grid), dim3(block), (size_t)(smem), (strm)>>>(__VA_ARGS__);        \
    } while (0)
#define TB_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/tobias_b200.h"

typedef long long i64;
typedef unsigned long long u64;
typedef unsigned int u32;

// A growable device allocation owned by the context.
struct tb_dbuf {
    void *p = nullptr;
    size_t cap = 0;
    template <class T> T *as() const { return (T *)p; }
};

// Bias model of one strand, device resident.
struct tb_model {
    int is_dwm = -1;
    int L = 0;
    tb_dbuf table;       // DWM: U[2 models][L][4][L][4] doubles; PWM: pssm[5][L]
    tb_dbuf vtable;      // DWM: V[2 models][16 * (L / 2) + 4 rows][L][4] doubles = 2^(pair / single rows of U), see tb_score.cu
    std::vector<double> host_raw[4];
    int product_ok = 1;  // DWM: every product of table rows stays below 2^480 in magnitude (the product-form kernel multiplies two of them unsplit)
};

struct tb_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;      // device-to-host copies that overlap with later kernels (tb_correct_download_begin)
    cudaStream_t aux_stream = nullptr;       // small table uploads that must not wait for the compute stream (a pageable copy synchronises its stream first)
    cudaEvent_t ev_aux = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_copy = nullptr;
    cudaEvent_t ev_chunk[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // chunks of a run handed to the copy stream
    cudaEvent_t marks[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float stage_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float last_ms = 0.f;
    i64 launches = 0;
    char err[512] = {0};

    // genome
    int n_contigs = 0;
    std::vector<i64> contig_len, contig_off;   // linear (padded) offsets in bases
    i64 genome_bases = 0;
    tb_dbuf seq2;      // u32 words, 16 bases each
    tb_dbuf nmask;     // u32 words, 32 bases each
    tb_dbuf d_contig_off;

    // reads
    i64 n_reads = 0;
    tb_dbuf r_start, r_len, r_clip;   // i64 linear start, i32 reference length, i32 5' clip
    tb_dbuf r_flags;
    bool reads_sorted = false;        // records sorted by (contig, start): lets K2 narrow the record range of a region chunk
    i64 max_read_len = 0;             // longest reference span of a record (bounds the records a region chunk can see)
    tb_dbuf k2_sites;                 // K2: site list (region, strand, offset of the cut site)
    tb_dbuf k2_hist;                  // K2: joint chunk histograms (u32 [4 tensors][chunk pairs][4^ch][4^ch])
    tb_dbuf k2_wts;                   // K2: capped multiplicity of every site
    // asynchronous genome uploads (tb_genome_upload_async): copies + packing run on copy_stream; consumers on `stream` wait for the
    // events of the contigs they read (tb_genome_fence), so work on the first contigs overlaps with the transfer of the later ones
    tb_dbuf staging_up;
    std::vector<tb_dbuf> nrun_buf;     // N-run tables of tb_genome_upload_packed_runs, one per contig
    std::vector<cudaEvent_t> contig_ready;
    std::vector<char> contig_pending;

    // scratch / per-call device tables
    std::vector<i64> h64[4];              // host scratch of the entry points (kept: a repeated call touches warm pages)
    std::vector<int32_t> h32[2];
    tb_dbuf reg_a, reg_b, reg_c, reg_d, reg_e;
    tb_dbuf scratch0, scratch1, scratch2, scratch3, scratch4;
    tb_dbuf staging;   // device staging for uploads
    void *pinned = nullptr;
    size_t pinned_cap = 0;

    // models
    tb_model model[2];

    // state of the last tb_correct_run
    struct {
        i64 n_regions = 0, ext_total = 0, out_total = 0, n_windows = 0;
        int k_flank = 0, window = 0, L = 0;
        std::vector<i64> ext_start, ext_off, out_off, win_off;   // linear ext start, offsets
        std::vector<i64> h_ls, h_le, h_ext_end;                  // host scratch of a run, kept so that a second run touches warm pages
        std::vector<int32_t> h_tile_region, h_tile_start;
        tb_dbuf d_ext_start, d_ext_end, d_ext_off, d_out_off, d_win_off;
        tb_dbuf pile;      // u32 [2][ext_total]
        tb_dbuf biasraw;   // f64 [2][ext_total]
        tb_dbuf fit;       // f64 [2][n_windows][6]
        tb_dbuf fitg_x, fitg_y, fitg_st, fitg_sti, fitg_wq, fitg_lists, fitg_sw;   // K3b round form: transposed window data, fit state, work lists
        tb_dbuf tracks;    // f64 [4 tracks][2 strands][out_total]
        tb_dbuf prepost;   // f64 [2][3][5][L]
        tb_dbuf dl[4][2];  // packed tracks on their way to the host
        double *pp_pinned = nullptr, *pp_dst = nullptr;   // pinned staging of prepost, caller's destination (set by _begin)
        bool valid = false;
        // destination of the next run's tracks (tb_correct_set_sink): the run hands them out in chunks while it computes
        bool sink_on = false;
        int sink_split = 0, sink_f64 = 0;
        void *sink_ptr[4][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
        double *sink_pp = nullptr;
    } cor;

    // state of K4
    struct {
        i64 n_regions = 0, ext_total = 0, out_total = 0;
        std::vector<i64> ext_off, out_off;
        tb_dbuf d_ext_off;
        tb_dbuf signal;    // f32 [ext_total]
        tb_dbuf work0, work1;   // f64 [ext_total]
        tb_dbuf out;       // trimmed output
        int out_f64 = 0;
        int flank = -1;
        bool have_signal = false, have_out = false;
        void *sink = nullptr;   // destination of the next tb_score_run's output (tb_score_set_sink)
    } sc;

    int k3a_tmem_groups = 0;    // K3a: pair groups kept in tensor memory (0 = default 5; option "k3a_tmem_groups" = 1..5, comparison runs)
    int k3a_no_tmem = 0;        // K3a: 1 = all table rows from shared memory (comparison runs; option "k3a_no_tmem")
    int force_sum_form = 0;     // K3a: 1 = first-generation sum-form kernel (comparison runs; option "k3a_sum_form")
    void *nccl_comm = nullptr;
    int rank = 0, world = 1;
    std::map<std::string, long long> opts;      // tb_set_option: kernel-variant / chunking switches for comparison runs and tests
};

static inline long long tb_opt(const tb_ctx *ctx, const char *key, long long dflt) {
    auto it = ctx->opts.find(key);
    return it == ctx->opts.end() ? dflt : it->second;
}

int tb_fail(tb_ctx *ctx, int code, const char *fmt, ...);
int tb_reserve(tb_ctx *ctx, tb_dbuf &b, size_t bytes);
int tb_genome_fence(tb_ctx *ctx, int c_lo, int c_hi);      // `stream` waits for pending asynchronous uploads of contigs [c_lo, c_hi]
void tb_release(tb_dbuf &b);
int tb_pinned(tb_ctx *ctx, size_t bytes);
int tb_check(tb_ctx *ctx, cudaError_t e, const char *what);
// host clock at the phase borders of an entry point, to stderr (option "host_trace"; comparison runs)
struct tb_tracer {
    const char *name;
    bool on;
    std::chrono::steady_clock::time_point t0;
    tb_tracer(const tb_ctx *ctx, const char *name_) : name(name_), on(tb_opt(ctx, "host_trace", 0) != 0), t0(std::chrono::steady_clock::now()) {}
    void mark(const char *what) const {
        if (on) fprintf(stderr, "[%s] %-28s %9.3f ms\n", name, what,
                        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
    ~tb_tracer() { mark("return"); }
};

int tb_h2d(tb_ctx *ctx, tb_dbuf &dst, const void *src, size_t bytes);          // reserve + async copy on ctx->stream
int tb_h2d_aux(tb_ctx *ctx, tb_dbuf &dst, const void *src, size_t bytes);      // reserve + copy on the aux stream: the host does not wait for the compute stream
int tb_aux_join(tb_ctx *ctx);                                                  // the compute stream waits for the aux copies issued so far
int tb_d2h(tb_ctx *ctx, void *dst, const void *src, size_t bytes);            // sync copy
int tb_sync(tb_ctx *ctx, const char *what);
void tb_timer_start(tb_ctx *ctx);
void tb_timer_stop(tb_ctx *ctx);
void tb_mark(tb_ctx *ctx, int i);                 // record stage boundary i on the context's stream
void tb_marks_collect(tb_ctx *ctx, int n_marks);  // stage_ms[i] = time between marks i and i+1
void tb_nccl_teardown(tb_ctx *ctx);
// regions (contig-local) -> linear coordinates; validates order; returns TB_OK or error
int tb_linear_regions(tb_ctx *ctx, i64 n, const int32_t *contig, const int64_t *start, const int64_t *end,
                      std::vector<i64> &lstart, std::vector<i64> &lend, bool need_disjoint);

#define TB_CUDA(ctx, call)                                        \
    do {                                                          \
        int _rc = tb_check((ctx), (call), #call);                 \
        if (_rc != TB_OK) return _rc;                             \
    } while (0)
#define TB_TRY(call)                                              \
    do {                                                          \
        int _rc = (call);                                         \
        if (_rc != TB_OK) return _rc;                             \
    } while (0)

static inline i64 tb_div_up(i64 a, i64 b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------- device helpers
// Genome: base k of the linear coordinate space lives in seq2[k >> 4] bits 2*(k & 15); N flag in nmask[k >> 5] bit (k & 31).
struct tb_genome_view {
    const u32 *seq2;
    const u32 *nmask;
    i64 n_bases;
};

__device__ __forceinline__ int tb_base(const tb_genome_view &g, i64 k) {
    return (int)((g.seq2[k >> 4] >> (2 * (int)(k & 15))) & 3u);
}
__device__ __forceinline__ int tb_is_n(const tb_genome_view &g, i64 k) {
    return (int)((g.nmask[k >> 5] >> (int)(k & 31)) & 1u);
}
// nb (<= 32) bases starting at k as a packed 2-bit word (base i in bits 2i..2i+1)
__device__ __forceinline__ u64 tb_bases(const tb_genome_view &g, i64 k, int nb) {
    i64 w = k >> 4;
    int sh = 2 * (int)(k & 15);
    u64 lo = (u64)g.seq2[w] | ((u64)g.seq2[w + 1] << 32);
    u64 v = lo >> sh;
    if (sh) v |= (u64)g.seq2[w + 2] << (64 - sh);
    if (nb < 32) v &= (((u64)1 << (2 * nb)) - 1);
    return v;
}
// N flags of nb (<= 32) bases starting at k (bit i = base k+i is N)
__device__ __forceinline__ u32 tb_nflags(const tb_genome_view &g, i64 k, int nb) {
    i64 w = k >> 5;
    int sh = (int)(k & 31);
    u64 v = (u64)g.nmask[w] | ((u64)g.nmask[w + 1] << 32);
    u32 r = (u32)(v >> sh);
    if (nb < 32) r &= ((1u << nb) - 1u);
    return r;
}
// the same for a value shared by the 32 lanes of a warp (all of them must call it): every round probes 32 pivots at once, so the
// chain of dependent loads is log33(n) long instead of log2(n) (4 rounds instead of 18 for 150k regions)
__device__ __forceinline__ i64 tb_upper_bound_warp(const i64 *a, i64 n, i64 v) {
    const int lane = threadIdx.x & 31;
    i64 lo = 0, hi = n;                                   // a[i] <= v below lo, a[i] > v from hi on
    while (lo < hi) {
        const i64 step = (hi - lo + 32) / 33;
        const i64 idx = lo + (i64)(lane + 1) * step - 1;
        const bool gt = idx < hi ? a[idx] > v : true;
        const unsigned m = __ballot_sync(0xffffffffu, gt);
        const int first = m ? __ffs((int)m) - 1 : 32;
        const i64 nlo = first == 0 ? lo : lo + (i64)first * step;
        i64 nhi = lo + (i64)(first + 1) * step - 1;
        if (first == 32 || nhi > hi) nhi = hi;
        lo = nlo;
        hi = nhi;
    }
    return lo;
}
// Exclusive block-wide prefix sums of two fp64 sequences held in shared memory, in place: a[i] = sum of the inputs below i, a[n] / b[n]
// receive the totals.  Every thread of the block must call it; wtot: 64 doubles of shared scratch.  Ends with a barrier.
__device__ __forceinline__ void tb_block_scan2w(double *a, double *b, int n, double *wtot) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int per = (n + nt - 1) / nt;
    const int lo = min(n, tid * per), hi = min(n, lo + per);
    double sa = 0.0, sb = 0.0;
    for (int i = lo; i < hi; i++) {
        sa += a[i];
        sb += b[i];
    }
    double ia = sa, ib = sb;                                      // inclusive scan over the lanes
    for (int d = 1; d < 32; d <<= 1) {
        const double oa = __shfl_up_sync(0xffffffffu, ia, d), ob = __shfl_up_sync(0xffffffffu, ib, d);
        if (lane >= d) {
            ia += oa;
            ib += ob;
        }
    }
    if (lane == 31) {
        wtot[warp] = ia;
        wtot[32 + warp] = ib;
    }
    __syncthreads();
    double ra = ia - sa, rb = ib - sb;
    for (int ww = 0; ww < warp; ww++) {
        ra += wtot[ww];
        rb += wtot[32 + ww];
    }
    for (int i = lo; i < hi; i++) {
        const double va = a[i], vb = b[i];
        a[i] = ra;
        b[i] = rb;
        ra += va;
        rb += vb;
    }
    if ((hi == n && lo < n) || (n == 0 && tid == 0)) {
        a[n] = ra;
        b[n] = rb;
    }
    __syncthreads();
}

// first index i in [0, n) with a[i] > v  (a ascending)
__device__ __forceinline__ i64 tb_upper_bound(const i64 *a, i64 n, i64 v) {
    i64 lo = 0, hi = n;
    while (lo < hi) {
        i64 mid = (lo + hi) >> 1;
        if (a[mid] > v) hi = mid; else lo = mid + 1;
    }
    return lo;
}
