/*
 * cuda_emu.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A tiny single-header CUDA execution-model emulator so that the kernels in tobias_b200/csrc/*.cu can be
 * compiled with g++ (-DTB_EMU) and their LOGIC exercised by the CPU test-suite in a container without a GPU.
 * Every CUDA thread of a block is a ucontext fiber; __syncthreads / __syncwarp / warp shuffles / ballots are
 * real rendezvous points between the fibers, so kernels run unmodified (same source, same indexing, same
 * barriers).  Blocks run one after another on the calling thread, hence atomics are plain read-modify-writes.
 *
 * This is not a product path: tobias_b200/lib.py only ever loads the nvcc-built libtobias_b200.so and fails
 * loudly when it is missing.  The emulator library is built by tests/emu/build_emu.py into tests/emu/_build/
 * and is loaded explicitly by CPU-only tests.
 */
#pragma once
#include <time.h>
#include <ucontext.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

using std::max;
using std::min;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __shared__ static
#define __constant__ static
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct double2 { double x, y; };
struct float2 { float x, y; };
inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
struct float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct ulonglong2 { unsigned long long x, y; };

typedef int cudaError_t;
typedef void *cudaStream_t;
struct tb_emu_event { double t; };
typedef tb_emu_event *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyHostToHost };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaStreamNonBlocking = 1, cudaEventDefault = 0, cudaHostAllocDefault = 0, cudaEventDisableTiming = 2 };
struct cudaDeviceProp { int multiProcessorCount; char name[64]; size_t totalGlobalMem; int major, minor; };

namespace tb_emu {

struct Fiber {
    ucontext_t ctx;
    uint3 tid;
    int linear;
    bool done;
    char *stack;
};

struct State {
    ucontext_t sched;
    std::vector<Fiber> fibers;
    std::vector<char> stacks;
    Fiber *cur = nullptr;
    uint3 bid{0, 0, 0};
    dim3 bdim, gdim;
    int nthreads = 0;
    // block barrier
    int block_arrived = 0;
    unsigned long block_gen = 0;
    // warp barriers
    std::vector<int> warp_arrived;
    std::vector<unsigned long> warp_gen;
    std::vector<int> warp_alive;
    std::vector<unsigned long long> warp_slot;   // 32 slots per warp for shuffles / ballots
    int alive = 0;
    unsigned char *dyn_smem = nullptr;
    std::vector<unsigned char> dyn_buf;
    std::function<void()> body;
    long launches = 0;
};

inline State &S() {
    static State s;
    return s;
}

static const size_t STACK_BYTES = 128 * 1024;

inline void yield_() {
    State &s = S();
    swapcontext(&s.cur->ctx, &s.sched);
}

inline void trampoline() {
    State &s = S();
    s.body();
    s.cur->done = true;
    s.alive--;
    int w = s.cur->linear >> 5;
    s.warp_alive[w]--;
    // a thread that exits releases barriers the remaining threads may be waiting on
    if (s.alive > 0 && s.block_arrived == s.alive) { s.block_arrived = 0; s.block_gen++; }
    if (s.warp_alive[w] > 0 && s.warp_arrived[w] == s.warp_alive[w]) { s.warp_arrived[w] = 0; s.warp_gen[w]++; }
    swapcontext(&s.cur->ctx, &s.sched);
}

inline void block_barrier() {
    State &s = S();
    unsigned long g = s.block_gen;
    if (++s.block_arrived == s.alive) { s.block_arrived = 0; s.block_gen++; return; }
    while (s.block_gen == g) yield_();
}

inline void warp_barrier() {
    State &s = S();
    int w = s.cur->linear >> 5;
    unsigned long g = s.warp_gen[w];
    if (++s.warp_arrived[w] == s.warp_alive[w]) { s.warp_arrived[w] = 0; s.warp_gen[w]++; return; }
    while (s.warp_gen[w] == g) yield_();
}

inline void run_block() {
    State &s = S();
    int n = s.nthreads;
    int nw = (n + 31) / 32;
    s.block_arrived = 0;
    s.alive = n;
    s.warp_arrived.assign(nw, 0);
    s.warp_gen.assign(nw, 0);
    s.warp_alive.assign(nw, 0);
    s.warp_slot.assign((size_t)nw * 32, 0);
    for (int t = 0; t < n; t++) {
        Fiber &f = s.fibers[t];
        f.done = false;
        f.linear = t;
        f.tid.x = t % s.bdim.x;
        f.tid.y = (t / s.bdim.x) % s.bdim.y;
        f.tid.z = t / (s.bdim.x * s.bdim.y);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = STACK_BYTES;
        f.ctx.uc_link = &s.sched;
        makecontext(&f.ctx, (void (*)())trampoline, 0);
        s.warp_alive[t >> 5]++;
    }
    while (s.alive > 0) {
        for (int t = 0; t < n; t++) {
            Fiber &f = s.fibers[t];
            if (f.done) continue;
            s.cur = &f;
            swapcontext(&s.sched, &f.ctx);
        }
    }
    s.cur = nullptr;
}

template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem, F body) {
    State &s = S();
    s.launches++;
    s.bdim = block;
    s.gdim = grid;
    s.nthreads = (int)(block.x * block.y * block.z);
    if ((int)s.fibers.size() < s.nthreads) {
        size_t old = s.fibers.size();
        s.fibers.resize(s.nthreads);
        for (size_t i = old; i < s.fibers.size(); i++) s.fibers[i].stack = (char *)malloc(STACK_BYTES);
    }
    s.dyn_buf.assign(smem + 64, 0);
    s.dyn_smem = (unsigned char *)(((uintptr_t)s.dyn_buf.data() + 63) & ~(uintptr_t)63);
    s.body = body;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                s.bid.x = bx; s.bid.y = by; s.bid.z = bz;
                run_block();
            }
}

inline int lane() { return S().cur->linear & 31; }
inline unsigned long long &slot(int l) { State &s = S(); return s.warp_slot[(size_t)(s.cur->linear >> 5) * 32 + l]; }

template <class T>
inline T shfl_generic(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle of > 8 bytes");
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    slot(lane()) = raw;
    warp_barrier();
    unsigned long long got = slot(src_lane & 31);
    warp_barrier();
    T out;
    memcpy(&out, &got, sizeof(T));
    return out;
}

}  // namespace tb_emu

#define threadIdx (tb_emu::S().cur->tid)
#define blockIdx (tb_emu::S().bid)
#define blockDim (tb_emu::S().bdim)
#define gridDim (tb_emu::S().gdim)
#define warpSize 32

inline void __syncthreads() { tb_emu::block_barrier(); }
inline void __syncwarp(unsigned = 0xffffffffu) { tb_emu::warp_barrier(); }
inline void __threadfence() {}
inline void __threadfence_block() {}

template <class T> inline T __shfl_sync(unsigned, T v, int src, int = 32) { return tb_emu::shfl_generic(v, src); }
template <class T> inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
    int l = tb_emu::lane();
    int src = l + (int)d;
    T got = tb_emu::shfl_generic(v, src > 31 ? l : src);
    return src > 31 ? v : got;
}
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
    int l = tb_emu::lane();
    int src = l - (int)d;
    T got = tb_emu::shfl_generic(v, src < 0 ? l : src);
    return src < 0 ? v : got;
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) {
    return tb_emu::shfl_generic(v, tb_emu::lane() ^ m);
}
inline unsigned __ballot_sync(unsigned, int pred) {
    tb_emu::slot(tb_emu::lane()) = pred ? 1ull : 0ull;
    tb_emu::warp_barrier();
    unsigned r = 0;
    tb_emu::State &s = tb_emu::S();
    int w = s.cur->linear >> 5;
    int n = std::min(32, s.nthreads - w * 32);
    for (int l = 0; l < n; l++)
        if (!s.fibers[w * 32 + l].done && tb_emu::slot(l)) r |= 1u << l;
    tb_emu::warp_barrier();
    return r;
}
inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
inline int __all_sync(unsigned m, int p) {
    unsigned b = __ballot_sync(m, p);
    unsigned a = __ballot_sync(m, 1);
    return b == a;
}
inline unsigned __activemask() { return 0xffffffffu; }

inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
    unsigned long long v = ((unsigned long long)hi << 32) | lo;
    return (unsigned)(v >> (sh & 31));
}
template <class T> inline T __ldg(const T *p) { return *p; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __double2float_rn(double a) { return (float)a; }
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
inline long long __double_as_longlong(double d) { long long i; memcpy(&i, &d, 8); return i; }
inline double __longlong_as_double(long long i) { double d; memcpy(&d, &i, 8); return d; }
inline int __double2hiint(double d) { long long i; memcpy(&i, &d, 8); return (int)(i >> 32); }
inline int __double2loint(double d) { long long i; memcpy(&i, &d, 8); return (int)(i & 0xffffffffll); }
inline double __hiloint2double(int hi, int lo) {
    unsigned long long i = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo;
    double d; memcpy(&d, &i, 8); return d;
}

template <class T> inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
template <class T> inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> inline T atomicCAS(T *p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }
template <class T> inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
template <class T> inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }

/* ---- runtime API subset ---- */
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 16; return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    memset(p, 0, sizeof(*p));
    p->multiProcessorCount = 4;
    strcpy(p->name, "tb_emu (CPU kernel emulator)");
    p->totalGlobalMem = (size_t)8 << 30;
    p->major = 10;
    return cudaSuccess;
}
inline cudaError_t cudaMemGetInfo(size_t *f, size_t *t) { *f = (size_t)8 << 30; *t = (size_t)8 << 30; return cudaSuccess; }
inline cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { if (n) memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void *d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = 0) { if (n) memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
inline const char *cudaGetErrorString(cudaError_t) { return "tb_emu"; }
inline double tb_emu_now() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new tb_emu_event{0}; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = 0) { e->t = tb_emu_now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
template <class F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
