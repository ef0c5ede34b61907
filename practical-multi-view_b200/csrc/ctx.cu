// ctx.cu -- context lifetime for libpmv_cuda.so (see include/pmv_cuda.h).
#include "common.cuh"

extern "C" {

PMV_API const char *pmv_version(void) { return "pmv_cuda 0.1 (sm_100a)"; }

PMV_API pmv_ctx *pmv_create(int device)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) {
        cudaGetLastError();
        return nullptr;  // no CUDA device: there is no CPU fallback
    }
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    pmv_ctx *c = new pmv_ctx();
    c->device = device;
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return nullptr;
    }
    for (auto &e : c->ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    {   // blocks freed with cudaFreeAsync stay in the device's default pool (bundle-adjustment problems come and go)
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    c->stream = c->own_stream;
    return c;
}

PMV_API void pmv_destroy(pmv_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->copy_stream);
    pmv_comm_destroy(c);   // the communicator belongs to the context (no-op when none was made)
    for (auto &b : c->img) b.release();
    for (auto &b : c->pyr) b.release();
    for (auto &b : c->pts) b.release();
    for (auto &b : c->scratch) b.release();
    for (auto &b : c->pin) b.release();
    for (auto &e : c->ev) if (e) cudaEventDestroy(e);
    for (auto &e : c->chunk_ev) cudaEventDestroy(e);
    for (auto &r : c->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto &e : c->prof_pool) cudaEventDestroy(e);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    delete c;
}

PMV_API int pmv_set_stream(pmv_ctx *c, void *s)
{
    if (!c) return PMV_ERR_INVALID;
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return PMV_OK;
}

PMV_API int pmv_sync(pmv_ctx *c)
{
    if (!c) return PMV_ERR_INVALID;
    PMV_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return PMV_OK;
}

PMV_API const char *pmv_last_error(pmv_ctx *c) { return c ? c->err.c_str() : "null context"; }

PMV_API int pmv_profile_enable(pmv_ctx *c, int on)
{
    if (!c) return PMV_ERR_INVALID;
    c->prof_on = on != 0;
    return PMV_OK;
}

PMV_API int pmv_profile_collect(pmv_ctx *c, int n_phases, double *ms_sum, int *count)
{
    if (!c || n_phases <= 0 || !ms_sum || !count) return PMV_ERR_INVALID;
    for (int i = 0; i < n_phases; i++) { ms_sum[i] = 0; count[i] = 0; }
    PMV_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    for (auto &r : c->prof_recs) {
        float ms = 0.f;
        cudaEventSynchronize(r.b);
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess && r.phase < n_phases) {
            ms_sum[r.phase] += ms;
            count[r.phase]++;
        }
        c->prof_pool.push_back(r.a);
        c->prof_pool.push_back(r.b);
    }
    c->prof_recs.clear();
    return PMV_OK;
}

PMV_API uint64_t pmv_launch_count(pmv_ctx *c) { return c ? c->launches : 0; }

}  // extern "C"
