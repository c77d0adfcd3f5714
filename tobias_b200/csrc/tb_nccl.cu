// tb_nccl.cu -- the one exchange step of the path: sum of the bias-training count tensors and normalisation counters
// over the ranks of a one-process-per-GPU job (AtacBias.join / SequenceMatrix.add_counts,
// tobias/tools/atacorrect_functions.py:54-60, tobias/utils/sequences.pyx:35-42; driver loop tobias/tools/atacorrect.py:324-326).
// The buffers are < 1 MB and integer valued, so one ncclAllReduce(sum) over NVLink/NVSwitch is exact and latency bound.
// NCCL is bound at run time (dlopen of libnccl.so.2) so that the library shares the NCCL instance of the host process
// (e.g. the one PyTorch loaded) instead of pulling in a second copy.
#include "tb_common.cuh"

#ifndef TB_EMU
#include <dlfcn.h>

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt64 = 4, ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct tb_nccl_api {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static tb_nccl_api *tb_nccl_load(char *err, size_t errlen) {
    static tb_nccl_api api;
    if (api.handle) return &api;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) {
        snprintf(err, errlen, "cannot load NCCL: %s", dlerror());
        return nullptr;
    }
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
    api.AllReduce = (decltype(api.AllReduce))dlsym(api.handle, "ncclAllReduce");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy || !api.GetErrorString) {
        snprintf(err, errlen, "NCCL library lacks a required symbol");
        api.handle = nullptr;
        return nullptr;
    }
    return &api;
}

extern "C" int tb_nccl_unique_id(uint8_t id_out[128]) {
    char err[256];
    tb_nccl_api *api = tb_nccl_load(err, sizeof(err));
    if (!api) return tb_fail(nullptr, TB_ERR_NCCL, "%s", err);
    ncclUniqueId id;
    ncclResult_t rc = api->GetUniqueId(&id);
    if (rc != 0) return tb_fail(nullptr, TB_ERR_NCCL, "ncclGetUniqueId: %s", api->GetErrorString(rc));
    memcpy(id_out, id.internal, 128);
    return TB_OK;
}

extern "C" int tb_nccl_init(tb_ctx *ctx, const uint8_t nccl_unique_id[128], int rank, int world_size) {
    if (!ctx || !nccl_unique_id) return TB_ERR_ARG;
    if (world_size < 1 || rank < 0 || rank >= world_size) return tb_fail(ctx, TB_ERR_ARG, "tb_nccl_init: bad rank %d / %d", rank, world_size);
    char err[256];
    tb_nccl_api *api = tb_nccl_load(err, sizeof(err));
    if (!api) return tb_fail(ctx, TB_ERR_NCCL, "%s", err);
    TB_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(id.internal, nccl_unique_id, 128);
    ncclComm_t comm = nullptr;
    ncclResult_t rc = api->CommInitRank(&comm, world_size, id, rank);
    if (rc != 0) return tb_fail(ctx, TB_ERR_NCCL, "ncclCommInitRank: %s", api->GetErrorString(rc));
    ctx->nccl_comm = comm;
    ctx->rank = rank;
    ctx->world = world_size;
    return TB_OK;
}

static int tb_allreduce(tb_ctx *ctx, void *buf, i64 n, int dtype) {
    if (!ctx || n < 0 || (n > 0 && !buf)) return TB_ERR_ARG;
    if (ctx->world == 1 || n == 0) return TB_OK;
    if (!ctx->nccl_comm) return tb_fail(ctx, TB_ERR_STATE, "tb_allreduce: call tb_nccl_init first");
    char err[256];
    tb_nccl_api *api = tb_nccl_load(err, sizeof(err));
    if (!api) return tb_fail(ctx, TB_ERR_NCCL, "%s", err);
    size_t bytes = (size_t)n * 8;
    TB_TRY(tb_h2d(ctx, ctx->scratch2, buf, bytes));
    ncclResult_t rc = api->AllReduce(ctx->scratch2.p, ctx->scratch2.p, (size_t)n, dtype, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream);
    if (rc != 0) return tb_fail(ctx, TB_ERR_NCCL, "ncclAllReduce: %s", api->GetErrorString(rc));
    return tb_d2h(ctx, buf, ctx->scratch2.p, bytes);
}

extern "C" int tb_allreduce_i64(tb_ctx *ctx, int64_t *buf, int64_t n) { return tb_allreduce(ctx, buf, n, ncclInt64); }
extern "C" int tb_allreduce_f64(tb_ctx *ctx, double *buf, int64_t n) { return tb_allreduce(ctx, buf, n, ncclFloat64); }

void tb_nccl_teardown(tb_ctx *ctx) {
    if (ctx->nccl_comm) {
        char err[256];
        tb_nccl_api *api = tb_nccl_load(err, sizeof(err));
        if (api) api->CommDestroy((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
}

#else  // TB_EMU: the kernel emulator has no collective; single-rank calls are no-ops, multi-rank calls fail loudly

extern "C" int tb_nccl_unique_id(uint8_t id_out[128]) {
    memset(id_out, 0, 128);
    return tb_fail(nullptr, TB_ERR_NCCL, "kernel emulator build: NCCL is not available");
}
extern "C" int tb_nccl_init(tb_ctx *ctx, const uint8_t *, int, int) {
    return tb_fail(ctx, TB_ERR_NCCL, "kernel emulator build: NCCL is not available");
}
extern "C" int tb_allreduce_i64(tb_ctx *ctx, int64_t *, int64_t) {
    return ctx && ctx->world == 1 ? TB_OK : tb_fail(ctx, TB_ERR_NCCL, "kernel emulator build: NCCL is not available");
}
extern "C" int tb_allreduce_f64(tb_ctx *ctx, double *, int64_t) {
    return ctx && ctx->world == 1 ? TB_OK : tb_fail(ctx, TB_ERR_NCCL, "kernel emulator build: NCCL is not available");
}
void tb_nccl_teardown(tb_ctx *) {}
#endif
