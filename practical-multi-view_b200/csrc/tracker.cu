// tracker.cu -- the front end of OdometryPipeline::addFrame with everything resident on the device.
//
// Replaces, as ONE handle that lives across frames,
//   matcher->matchFeatures(prev, frame)            reference OdometryPipeline.cpp:336  (OpenCVLucasKanadeFM.cpp:5-32)
//   the status filter + Feature(int, int) truncation                                   (OpenCVLucasKanadeFM.cpp:23-29)
//   the ROI-grid re-extraction with Frame::hasNeighbor de-duplication                  (OdometryPipeline.cpp:343-370, Frame.cpp:3-12)
//   initialise()'s ROI-grid extraction on the first frame                              (OdometryPipeline.cpp:440-459)
// The reference rebuilds both image pyramids inside every calcOpticalFlowPyrLK call (frame k is reduced twice: as
// `next`, then as `prev`) and keeps the tracks in host hash maps.  Here every frame is uploaded once, its pyramid AND
// its Scharr planes are built once by the fused level kernel (pyramid.cu) into one of two ping-pong sets, the track
// list stays on the device, and a frame costs one small download (count + the surviving features).
//
// Reference quirks that are reproduced on purpose (SURVEY Appx D): re-extraction runs on the PREVIOUS frame
// (frames.size() - 1 before the push_back), and hasNeighbor compares ROI-LOCAL candidate coordinates with the frame's
// global feature coordinates before the ROI offset is added.
#include "common.cuh"

struct pmv_tracker {
    pmv_ctx *ctx = nullptr;
    int rows = 0, cols = 0, win_w = 0, win_h = 0, max_level = 0;
    int cap = 0;                       // feature capacity
    int min_tracked = 400, tol = 150, grid = 255, nb_dist = 5;
    double quality = 0.01, min_dist = 5.0;
    DevBuf pyr[2], der[2];
    unsigned long long dsig[2] = {0, 0};
    PyrSet set[2];
    DerivSet dv[2];
    int cur = 0;                       // set holding the latest frame
    int frames = 0;
    int n_feat = 0;                    // features of the latest frame
    DevBuf feat[2];                    // float2 (column, row), integer-valued: the points handed to the tracker kernel
    DevBuf prev_idx;                   // int: index of each feature in the previous frame's list, -1 = newly extracted
    DevBuf nxy, st, err;               // raw tracker outputs
    DevBuf misc;                       // [0] feature count
    DevBuf raw;                        // uploaded frame, rows pitched to 16 B (the fused level kernel reads it through TMA)
    PinBuf pin;                        // count + feature download
    PinBuf pin_img;                    // staging of the caller's (pageable, any step) frame: one contiguous DMA per frame
    int raw_pitch = 0;
};

namespace {

// Stable compaction of the tracked points: status == 1 -> (int)x, (int)y (Feature(int, int), truncation toward zero),
// kept in the order of the previous list.  One CTA (a frame has hundreds to a few thousand features).
__global__ void __launch_bounds__(1024)
track_compact_kernel(const float2 *__restrict__ nxy, const uint8_t *__restrict__ st, int n, float2 *__restrict__ out,
                     int *__restrict__ prev_idx, int *__restrict__ count)
{
    __shared__ int warp_sum[32];
    __shared__ int base;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += 1024) {
        const int i = i0 + t;
        const bool keep = i < n && st[i] == 1;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_sum[w] = __popc(m);
        __syncthreads();
        int off = base;
        for (int k = 0; k < w; k++) off += warp_sum[k];
        if (keep) {
            const int o = off + __popc(m & ((1u << lane) - 1));
            const float2 p = nxy[i];
            out[o] = make_float2((float)(int)p.x, (float)(int)p.y);
            prev_idx[o] = i;
        }
        __syncthreads();
        if (t == 0) { int s = 0; for (int k = 0; k < 32; k++) s += warp_sum[k]; base += s; }
        __syncthreads();
    }
    if (t == 0) *count = base;
}

// Append the corners of one ROI to the feature list.  dedup != 0: Frame::hasNeighbor -- a candidate is dropped when any
// feature already in the list (tracked ones and the candidates accepted before it) lies at Chebyshev distance < dist
// of its ROI-LOCAL coordinates; the ROI offset is added afterwards.  Greedy and sequential like the reference loop.
__global__ void __launch_bounds__(256)
append_roi_kernel(const float2 *__restrict__ cand, int n_cand, int off_x, int off_y, int dedup, int dist, float2 *__restrict__ feat,
                  int *__restrict__ prev_idx, int *__restrict__ count, int cap)
{
    __shared__ int n_s;
    if (threadIdx.x == 0) n_s = *count;
    __syncthreads();
    for (int c = 0; c < n_cand; c++) {
        const int n = n_s;
        const int cx = (int)cand[c].x, cy = (int)cand[c].y;
        int hit = 0;
        if (dedup) {
            for (int i = threadIdx.x; i < n; i += 256) {
                const float2 f = feat[i];
                const int dx = abs(cx - (int)f.x), dy = abs(cy - (int)f.y);
                hit |= (dx > dy ? dx : dy) < dist;
            }
        }
        hit = __syncthreads_or(hit);
        if (!hit && threadIdx.x == 0 && n < cap) {
            feat[n] = make_float2((float)(cx + off_x), (float)(cy + off_y));
            prev_idx[n] = -1;
            n_s = n + 1;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = n_s;
}

int upload_and_build(pmv_tracker *T, int which, const uint8_t *frame, int step, cudaStream_t s)
{
    pmv_ctx *ctx = T->ctx;
    // rows are repacked to a 16-byte pitch in pinned memory (a 2-D copy from pageable memory is staged row by row by
    // the driver), then ONE contiguous DMA; the previous frame's DMA has completed (every call ends with a sync)
    uint8_t *h = T->pin_img.as<uint8_t>();
    for (int y = 0; y < T->rows; y++) memcpy(h + (size_t)y * T->raw_pitch, frame + (size_t)y * step, T->cols);
    PMV_CUDA_TRY(ctx, cudaMemcpyAsync(T->raw.p, h, (size_t)T->raw_pitch * T->rows, cudaMemcpyHostToDevice, s));
    ProfScope ps(ctx, PMV_PHASE_PYRAMID, s);
    // the fused pass reads the upload through TMA and writes the bordered level-0 copy, the levels, their borders and
    // the Scharr planes this frame needs when it becomes `prev` on the next call
    return pmv_internal_pyr_run(ctx, T->set[which], 1, T->raw.as<uint8_t>(), nullptr, T->raw_pitch, (size_t)T->raw_pitch * T->rows,
                                &T->dv[which], 1, s);
}

// ROI-grid extraction on the resident level 0 of set `which` (OdometryPipeline::getGridROI order: rows, then columns)
int extract_grid(pmv_tracker *T, int which, int per_roi, int dedup, cudaStream_t s)
{
    pmv_ctx *ctx = T->ctx;
    const PyrLevel &l0 = T->set[which].lv[0];
    float2 *feat = T->feat[T->cur].as<float2>();
    for (int r = 0; r < T->rows; r += T->grid) {
        for (int c = 0; c < T->cols; c += T->grid) {
            const int rw = T->cols - c < T->grid ? T->cols - c : T->grid, rh = T->rows - r < T->grid ? T->rows - r : T->grid;
            float *d_xy = nullptr, *d_sc = nullptr;
            int n = 0;
            int rc = pmv_internal_gftt_device(ctx, l0.ptr, l0.pitch, T->rows, T->cols, c, r, rw, rh, per_roi, T->quality, T->min_dist,
                                              &d_xy, &d_sc, &n);
            if (rc) return rc;
            if (n <= 0) continue;
            // GridSection(x = c / grid, y = r / grid); offset = x * grid_size[1], y * grid_size[0]
            append_roi_kernel<<<1, 256, 0, s>>>(reinterpret_cast<const float2 *>(d_xy), n, (c / T->grid) * T->grid, (r / T->grid) * T->grid,
                                                dedup, T->nb_dist, feat, T->prev_idx.as<int>(), T->misc.as<int>(), T->cap);
            PMV_LAUNCH_CHECK(ctx, "append_roi_kernel");
        }
    }
    return PMV_OK;
}

int n_rois(const pmv_tracker *T) { return ((T->rows + T->grid - 1) / T->grid) * ((T->cols + T->grid - 1) / T->grid); }

}  // namespace

extern "C" {

PMV_API pmv_tracker *pmv_tracker_create(pmv_ctx *ctx, int rows, int cols, int win_w, int win_h, int max_level, int capacity,
                                        int min_tracked, int tracked_tol, int grid, double quality, double min_dist, int neighbor_dist)
{
    if (!ctx) return nullptr;
    if (rows <= 0 || cols <= 0 || win_w <= 2 || win_h <= 2 || win_w * win_h > 1024 || max_level < 0 || max_level >= PMV_MAX_PYR_LEVELS ||
        capacity <= 0 || grid <= 0 || quality <= 0 || min_dist < 0 || min_tracked <= 0) {
        ctx->fail(PMV_ERR_INVALID, "pmv_tracker_create: bad argument");
        return nullptr;
    }
    cudaSetDevice(ctx->device);
    pmv_tracker *T = new pmv_tracker;
    T->ctx = ctx; T->rows = rows; T->cols = cols; T->win_w = win_w; T->win_h = win_h; T->max_level = max_level;
    T->cap = capacity; T->min_tracked = min_tracked; T->tol = tracked_tol; T->grid = grid; T->quality = quality; T->min_dist = min_dist;
    T->nb_dist = neighbor_dist;
    const int border = pmv_internal_lk_border(win_w, win_h);
    int rc = 0;
    for (int k = 0; k < 2 && !rc; k++) {
        rc = pmv_internal_pyr_plan_buf(ctx, &T->pyr[k], 1, rows, cols, border, win_w, win_h, max_level, &T->set[k]);
        if (!rc) rc = pmv_internal_deriv_plan_buf(ctx, &T->der[k], &T->dsig[k], T->set[k], 1, &T->dv[k], ctx->stream);
    }
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < 2 && e == cudaSuccess; k++) e = T->feat[k].reserve((size_t)capacity * 8);
    if (e == cudaSuccess) e = T->prev_idx.reserve((size_t)capacity * 4);
    if (e == cudaSuccess) e = T->nxy.reserve((size_t)capacity * 8);
    if (e == cudaSuccess) e = T->st.reserve((size_t)capacity);
    if (e == cudaSuccess) e = T->err.reserve((size_t)capacity * 4);
    if (e == cudaSuccess) e = T->misc.reserve(64);
    if (e == cudaSuccess) e = T->pin.reserve(64 + (size_t)capacity * 12);
    T->raw_pitch = align_up(cols, 16);
    if (e == cudaSuccess) e = T->raw.reserve((size_t)T->raw_pitch * rows + 256);
    if (e == cudaSuccess) e = T->pin_img.reserve((size_t)T->raw_pitch * rows + 256);
    if (rc || e != cudaSuccess) {
        if (!rc) ctx->fail(PMV_ERR_NOMEM, "pmv_tracker_create: buffers", e);
        pmv_tracker_destroy(T);
        return nullptr;
    }
    return T;
}

PMV_API void pmv_tracker_destroy(pmv_tracker *T)
{
    if (!T) return;
    cudaSetDevice(T->ctx->device);
    cudaStreamSynchronize(T->ctx->stream);
    for (int k = 0; k < 2; k++) { T->pyr[k].release(); T->der[k].release(); T->feat[k].release(); }
    T->prev_idx.release(); T->nxy.release(); T->st.release(); T->err.release(); T->misc.release(); T->pin.release(); T->raw.release(); T->pin_img.release();
    delete T;
}

// initialise() on one frame: upload, pyramid + derivative planes, int(min_tracked / n_roi) corners per ROI (no de-dup).
PMV_API int pmv_tracker_init(pmv_tracker *T, const uint8_t *frame, int step, int *n_features)
{
    if (!T) return PMV_ERR_INVALID;
    pmv_ctx *ctx = T->ctx;
    if (!frame || step < T->cols) return ctx->fail(PMV_ERR_INVALID, "pmv_tracker_init: bad argument");
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    T->cur = 0; T->frames = 0; T->n_feat = 0;
    int rc = upload_and_build(T, 0, frame, step, s);
    if (rc) return rc;
    PMV_CUDA_TRY(ctx, cudaMemsetAsync(T->misc.p, 0, 16, s));
    rc = extract_grid(T, 0, T->min_tracked / n_rois(T), 0, s);
    if (rc) return rc;
    int *h = T->pin.as<int>();
    PMV_CUDA_TRY(ctx, cudaMemcpyAsync(h, T->misc.p, 4, cudaMemcpyDeviceToHost, s));
    PMV_CUDA_TRY(ctx, cudaStreamSynchronize(s));
    T->n_feat = h[0];
    T->frames = 1;
    if (n_features) *n_features = T->n_feat;
    return PMV_OK;
}

// addFrame(): one upload, one pyramid build, LK from the resident features of the previous frame, status filter and
// truncation on the device, ROI-grid re-extraction (on the PREVIOUS frame, like the reference) when fewer than
// tracked_tol survive.  xy / prev_index (optional, capacity entries): the new frame's features and where each came
// from in the previous frame's list (-1 = newly extracted).
PMV_API int pmv_tracker_add_frame(pmv_tracker *T, const uint8_t *frame, int step, int *n_tracked, int *n_features, int *extracted,
                                  int32_t *xy, int32_t *prev_index, int capacity)
{
    if (!T) return PMV_ERR_INVALID;
    pmv_ctx *ctx = T->ctx;
    if (!frame || step < T->cols) return ctx->fail(PMV_ERR_INVALID, "pmv_tracker_add_frame: bad argument");
    if (T->frames < 1) return ctx->fail(PMV_ERR_INVALID, "pmv_tracker_add_frame: call pmv_tracker_init first");
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    const int prev = T->cur, next = 1 - T->cur;
    int rc = upload_and_build(T, next, frame, step, s);
    if (rc) return rc;
    const int n = T->n_feat;
    float2 *fprev = T->feat[prev].as<float2>(), *fnext = T->feat[next].as<float2>();
    int *d_count = T->misc.as<int>();
    if (n > 0) {
        ProfScope pl(ctx, PMV_PHASE_LK, s);
        rc = pmv_internal_lk_launch(ctx, T->set[prev], T->set[next], T->dv[prev], 1, reinterpret_cast<const float *>(fprev), n, T->win_w,
                                    T->win_h, 30, 0.01, 0, 1e-4, T->nxy.as<float>(), T->st.as<uint8_t>(), T->err.as<float>(), s);
        if (rc) return rc;
    }
    track_compact_kernel<<<1, 1024, 0, s>>>(T->nxy.as<float2>(), T->st.as<uint8_t>(), n, fnext, T->prev_idx.as<int>(), d_count);
    PMV_LAUNCH_CHECK(ctx, "track_compact_kernel");
    // one download: the count and (at most n) surviving features with their previous indices
    int *h = T->pin.as<int>();
    float *hf = reinterpret_cast<float *>(h + 16);
    int *hp = h + 16 + 2 * T->cap;
    PMV_CUDA_TRY(ctx, cudaMemcpyAsync(h, d_count, 4, cudaMemcpyDeviceToHost, s));
    if (n > 0 && (xy || prev_index)) {
        PMV_CUDA_TRY(ctx, cudaMemcpyAsync(hf, fnext, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
        PMV_CUDA_TRY(ctx, cudaMemcpyAsync(hp, T->prev_idx.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    }
    PMV_CUDA_TRY(ctx, cudaStreamSynchronize(s));
    const int tracked = h[0];
    int total = tracked, did = 0;
    if (tracked < T->tol) {
        did = 1;
        T->cur = next;   // extract_grid appends to feat[cur]
        const int per_roi = (T->min_tracked + n_rois(T) - 1) / n_rois(T);   // std::ceil(min_tracked / roi.size())
        rc = extract_grid(T, prev, per_roi, 1, s);
        if (rc) { T->cur = prev; return rc; }
        PMV_CUDA_TRY(ctx, cudaMemcpyAsync(h, d_count, 4, cudaMemcpyDeviceToHost, s));
        PMV_CUDA_TRY(ctx, cudaStreamSynchronize(s));
        total = h[0];
        if (total > 0 && (xy || prev_index)) {
            PMV_CUDA_TRY(ctx, cudaMemcpyAsync(hf, fnext, (size_t)total * 8, cudaMemcpyDeviceToHost, s));
            PMV_CUDA_TRY(ctx, cudaMemcpyAsync(hp, T->prev_idx.p, (size_t)total * 4, cudaMemcpyDeviceToHost, s));
            PMV_CUDA_TRY(ctx, cudaStreamSynchronize(s));
        }
    }
    T->cur = next;
    T->n_feat = total;
    T->frames++;
    const int m = total < capacity ? total : capacity;
    if (xy) for (int i = 0; i < m; i++) { xy[2 * i] = (int32_t)hf[2 * i]; xy[2 * i + 1] = (int32_t)hf[2 * i + 1]; }
    if (prev_index) for (int i = 0; i < m; i++) prev_index[i] = hp[i];
    if (n_tracked) *n_tracked = tracked;
    if (n_features) *n_features = total;
    if (extracted) *extracted = did;
    return PMV_OK;
}

// Features of the latest frame (column, row) -- e.g. after pmv_tracker_init.
PMV_API int pmv_tracker_features(pmv_tracker *T, int32_t *xy, int capacity, int *n)
{
    if (!T) return PMV_ERR_INVALID;
    pmv_ctx *ctx = T->ctx;
    if (!n) return ctx->fail(PMV_ERR_INVALID, "pmv_tracker_features: null count");
    cudaSetDevice(ctx->device);
    const int m = T->n_feat < capacity ? T->n_feat : capacity;
    if (m > 0 && xy) {
        float *hf = reinterpret_cast<float *>(T->pin.as<int>() + 16);
        PMV_CUDA_TRY(ctx, cudaMemcpyAsync(hf, T->feat[T->cur].p, (size_t)m * 8, cudaMemcpyDeviceToHost, ctx->stream));
        PMV_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < m; i++) { xy[2 * i] = (int32_t)hf[2 * i]; xy[2 * i + 1] = (int32_t)hf[2 * i + 1]; }
    }
    *n = T->n_feat;
    return PMV_OK;
}

}  // extern "C"
