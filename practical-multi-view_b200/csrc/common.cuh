// common.cuh -- shared plumbing for libpmv_cuda.so (context, error handling, device helpers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/pmv_cuda.h"

// ------------------------------------------------------------------ device buffers -----
// Grow-only device buffer owned by a context (no per-call cudaMalloc on the hot path).
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct HostBuf {  // grow-only pageable host scratch that outlives a call (fresh malloc blocks of tens of MB fault their
                  // pages in again on every call: half of the run-construction time of a 1 M-point BA problem)
    void *p = nullptr;
    size_t cap = 0;
    void *reserve(size_t bytes)
    {
        if (bytes > cap) { free(p); p = malloc(bytes + 64); cap = p ? bytes + 64 : 0; }
        return p;
    }
    ~HostBuf() { free(p); }
};

struct PinBuf {  // grow-only pinned host staging buffer
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMallocHost(&p, bytes + 256);
        if (e == cudaSuccess) cap = bytes + 256;
        return e;
    }
    void release()
    {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// ------------------------------------------------------------------ pyramid layout ------
// Level l of a batch of B images lives at  ptr + b*img_stride + y*pitch + x  (u8), where ptr
// addresses pixel (0,0) of the INTERIOR.  Every context-owned level (level 0 included: caller
// images are imported) carries a BORDER_REFLECT_101 border of at least `border` pixels on all
// four sides, so the LK kernel never reflects coordinates: reads at x in [-border, cols+border)
// are plain loads.  ptr is 16 B aligned and pitch is a multiple of 128 B (vector loads / TMA).
struct PyrLevel {
    const uint8_t *ptr;
    int rows, cols, pitch;
    size_t img_stride;
    int border;  // reflect-101 border filled on all four sides (pixels)
    int bxl;     // allocated bytes left of interior column 0 (>= border + 4, multiple of 16)
};

// Scharr derivative of one level for a batch of images, as OpenCV's derivative pyramid keeps it: one packed
// word per pixel (dI/dx in the low, dI/dy in the high 16 bits), ZERO outside the image for `border` pixels on
// all sides.  ptr = pixel (0, 0) of image 0; pitch / img_stride in words; rows 128 B aligned.
struct DerivLevel {
    const int *ptr;
    int pitch;
    size_t img_stride;
};
struct DerivSet {
    DerivLevel lv[PMV_MAX_PYR_LEVELS];
};

struct PyrSet {  // all levels of one image batch
    PyrLevel lv[PMV_MAX_PYR_LEVELS];
    int top = 0;  // effective max level
};

enum { PMV_ATTR_CHOL_SMALL = 0, PMV_ATTR_CHOL_BACKSUB, PMV_ATTR_CHOL_BAND, PMV_ATTR_WIN_SCHUR, PMV_ATTR_CHOL_SPIKE, PMV_ATTR_BS_SCAN, PMV_ATTR_BS_SCAN_FAST, PMV_ATTR_LK_BASE /* + KPIX (<= 32) */ };

static inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

// ------------------------------------------------------------------ context -------------
struct pmv_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;      // stream work is enqueued on
    cudaStream_t own_stream = nullptr;  // created by pmv_create
    cudaStream_t copy_stream = nullptr; // second stream for chunked upload overlap
    cudaStream_t d2h_stream = nullptr;  // third stream: chunked result download (created on first use)
    cudaStream_t aux_stream = nullptr;  // second compute stream of the chunked host paths (created on first use)
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<cudaEvent_t> chunk_ev;  // upload/compute hand-off events of the chunked host paths
    std::string err;
    uint64_t launches = 0;
    // cudaFuncSetAttribute is per DEVICE: remembered per context (= per device), never per process.
    // One bit per kernel family (PMV_ATTR_*); attr_first() is true the first time a bit is asked for.
    uint64_t attr_done = 0;
    bool attr_first(int bit) { const uint64_t m = 1ull << bit; if (attr_done & m) return false; attr_done |= m; return true; }
    void *nccl_comm = nullptr;          // ncclComm_t of the sharded bundle adjuster (ba_nccl.cu)
    int nranks = 1, rank = 0;

    // optional per-phase CUDA-event timing (bench roofline); see pmv_profile_*
    bool prof_on = false;
    struct ProfRec { int phase; cudaEvent_t a, b; };
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
    cudaEvent_t prof_event()
    {
        cudaEvent_t e = nullptr;
        if (!prof_pool.empty()) { e = prof_pool.back(); prof_pool.pop_back(); }
        else cudaEventCreate(&e);
        return e;
    }

    // workspaces
    DevBuf img[2];    // uploaded level-0 images (prev / next)
    DevBuf pyr[2];    // reduced levels (prev / next)
    DevBuf deriv;     // Scharr derivative levels of the prev images (LK)
    unsigned long long deriv_sig = 0;   // geometry the zero borders of `deriv` were written for
    DevBuf pts[4];    // prev_xy, next_xy, status, err
    DevBuf scratch[8];
    PinBuf pin[4];
    HostBuf host[2];   // BA problem creation: tuple keys

    int last_code = 0;   // status of the last failure (entry points that return a handle report it through this)
    int fail(int code, const char *what, cudaError_t e = cudaSuccess)
    {
        last_code = code;
        char b[512];
        if (e != cudaSuccess)
            snprintf(b, sizeof b, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
        else
            snprintf(b, sizeof b, "%s", what);
        err = b;
        return code;
    }
};

#define PMV_CUDA_TRY(ctx, expr)                                                     \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) return (ctx)->fail(PMV_ERR_CUDA, #expr, _e);         \
    } while (0)

#define PMV_LAUNCH_CHECK(ctx, name)                                                 \
    do {                                                                            \
        (ctx)->launches++;                                                          \
        cudaError_t _e = cudaGetLastError();                                        \
        if (_e != cudaSuccess) return (ctx)->fail(PMV_ERR_CUDA, "launch " name, _e);\
    } while (0)

// RAII phase timer: records an event pair on `s` around a group of launches when profiling is on.
struct ProfScope {
    pmv_ctx *c; cudaStream_t s; int phase; cudaEvent_t a = nullptr;
    ProfScope(pmv_ctx *c_, int phase_, cudaStream_t s_) : c(c_), s(s_), phase(phase_)
    {
        if (c && c->prof_on) { a = c->prof_event(); cudaEventRecord(a, s); }
    }
    ~ProfScope()
    {
        if (a) { cudaEvent_t b = c->prof_event(); cudaEventRecord(b, s); c->prof_recs.push_back({phase, a, b}); }
    }
};

// ------------------------------------------------------------------ device helpers ------
__host__ __device__ __forceinline__ int reflect101(int p, int len)
{
    // BORDER_REFLECT_101  gfedcb|abcdefgh|gfedcba ; loop form is safe for any overshoot
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

// internal cross-file entry points (defined in pyramid.cu)
// Plan bordered storage for levels 0..L of `batch` images in ctx->pyr[which].
int pmv_internal_pyr_plan(pmv_ctx *ctx, int which, int batch, int rows, int cols, int border,
                          int win_w, int win_h, int max_level, PyrSet *out);
// the same in a buffer the caller owns (pmv_tracker keeps two pyramids alive between calls)
int pmv_internal_pyr_plan_buf(pmv_ctx *ctx, DevBuf *buf, int batch, int rows, int cols, int border,
                              int win_w, int win_h, int max_level, PyrSet *out);
// Import level 0 from device memory (any pitch) / expect it already copied into the interior
// (src == nullptr), fill its border, then build levels 1..top with their borders.
int pmv_internal_pyr_run(pmv_ctx *ctx, const PyrSet &set, int batch, const uint8_t *d_src, const uint8_t *d_src2,
                         int src_pitch, size_t src_stride, const DerivSet *dv, int n_deriv, cudaStream_t s);
// Plan (and zero, when the geometry changed) derivative storage matching the levels of `set`; the planes of the
// first n_deriv images are written by pmv_internal_pyr_run in the same pass that builds the levels.
int pmv_internal_deriv_plan(pmv_ctx *ctx, const PyrSet &set, int batch, DerivSet *out, cudaStream_t s);
int pmv_internal_deriv_plan_buf(pmv_ctx *ctx, DevBuf *buf, unsigned long long *sig, const PyrSet &set, int batch, DerivSet *out,
                                cudaStream_t s);
// internal cross-file entry points (defined in lk.cu): border the tracker needs around every level, and the tracking
// launch alone on pyramids / derivative planes that are already built (device pointers, asynchronous on s)
int pmv_internal_lk_border(int win_w, int win_h);
int pmv_internal_lk_launch(pmv_ctx *ctx, const PyrSet &sp, const PyrSet &sn, const DerivSet &dv, int batch,
                           const float *d_prev_xy, int n, int win_w, int win_h, int max_count, double eps, int flags,
                           double min_eig_thr, float *d_next_xy, uint8_t *d_status, float *d_err, cudaStream_t s);
// defined in corners.cu: goodFeaturesToTrack of an ROI of an image that is ALREADY on the device (bordered or not:
// reads outside the parent are reflected).  Results stay on the device (xy: n x 2 float ROI-local, score), *n on the host.
int pmv_internal_gftt_device(pmv_ctx *ctx, const uint8_t *d_img, int pitch, int full_rows, int full_cols, int rx, int ry, int rw,
                             int rh, int max_corners, double quality, double min_dist, float **d_xy, float **d_score, int *n);
