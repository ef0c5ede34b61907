// ba_nccl.cu -- the one exchange step of the hot path: point-sharded bundle adjustment sums each
// rank's partial reduced camera system [S | rhs], camera blocks and LM scalars with
// ncclAllReduce(ncclDouble) over NVLink.  NCCL is bound at run time with dlopen so the
// library loads on hosts without NCCL and shares the process-wide libnccl.so.2 torch already mapped.
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

namespace {

struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi &api()
{
    static NcclApi a;
    if (a.h) return a;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        a.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (a.h) break;
    }
    if (!a.h) return a;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.h, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.h, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.h, "ncclCommDestroy");
    a.AllReduce = (decltype(a.AllReduce))dlsym(a.h, "ncclAllReduce");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.h, "ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.GetErrorString;
    return a;
}

int nccl_fail(pmv_ctx *ctx, const char *what, ncclResult_t r)
{
    char b[256];
    snprintf(b, sizeof b, "%s: %s", what, api().GetErrorString ? api().GetErrorString(r) : "nccl error");
    return ctx->fail(PMV_ERR_NCCL, b);
}

}  // namespace

int pmv_internal_ba_allreduce(pmv_ctx *ctx, const double *send, double *recv, size_t count, int op_max, cudaStream_t s)
{
    if (!ctx->nccl_comm) return ctx->fail(PMV_ERR_NCCL, "sharded solve without pmv_comm_init");
    ncclResult_t r = api().AllReduce(send, recv, count, ncclDouble, op_max ? ncclMax : ncclSum,
                                     (ncclComm_t)ctx->nccl_comm, s);
    if (r != ncclSuccess) return nccl_fail(ctx, "ncclAllReduce", r);
    ctx->launches++;
    return PMV_OK;
}

extern "C" {

PMV_API int pmv_comm_unique_id(char id[128])
{
    if (!id || !api().ok) return PMV_ERR_NCCL;
    ncclUniqueId u;
    if (api().GetUniqueId(&u) != ncclSuccess) return PMV_ERR_NCCL;
    memcpy(id, u.internal, 128);
    return PMV_OK;
}

PMV_API int pmv_comm_init(pmv_ctx *ctx, int nranks, int rank, const char id[128])
{
    if (!ctx) return PMV_ERR_INVALID;
    if (!id || nranks < 1 || rank < 0 || rank >= nranks) return ctx->fail(PMV_ERR_INVALID, "pmv_comm_init: bad argument");
    if (!api().ok) return ctx->fail(PMV_ERR_NCCL, "libnccl.so.2 could not be loaded");
    cudaSetDevice(ctx->device);
    ncclUniqueId u;
    memcpy(u.internal, id, 128);
    ncclComm_t c = nullptr;
    ncclResult_t r = api().CommInitRank(&c, nranks, u, rank);
    if (r != ncclSuccess) return nccl_fail(ctx, "ncclCommInitRank", r);
    ctx->nccl_comm = c;
    ctx->nranks = nranks;
    ctx->rank = rank;
    return PMV_OK;
}

PMV_API int pmv_comm_destroy(pmv_ctx *ctx)
{
    if (!ctx) return PMV_ERR_INVALID;
    if (ctx->nccl_comm && api().ok) api().CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    ctx->nranks = 1;
    ctx->rank = 0;
    return PMV_OK;
}

}  // extern "C"
