// fast.cu -- K7: FAST-9/16 corner detector with score and 3x3 non-max suppression.
//
// Replaces cv::FAST(src.bw, kp, 10, true) at reference OpenCVFASTFeatureExtractor.cpp:8
// (TYPE_9_16; SURVEY Appx B.4).  Integer arithmetic, bit-exact: segment test on the radius-3
// Bresenham circle, score = largest threshold for which the pixel is still a corner
// (cornerScore<16>), NMS keeps a pixel iff its score beats all 8 neighbours; rows/cols 3..dim-4.
// The adapter keeps the first `max` keypoints in raster order, so the unordered atomic
// compaction is followed by the 128-bit bitonic sort on ~index.
//
// One CTA = 64x16 output tile from a (64+8)x(16+8) u8 tile staged in shared memory; the score
// tile (halo 1) lives in shared memory too, so the image is read once (1 B/px, HBM bound).
#include "common.cuh"
#include "sort.cuh"

namespace {

constexpr int FT_W = 64, FT_H = 16;
constexpr int FS_W = FT_W + 8, FS_H = FT_H + 8, FS_P = 72;
constexpr int FC_W = FT_W + 2, FC_H = FT_H + 2;

// radius-3 Bresenham circle; constexpr so that the staged-tile offsets fold into the load instructions
__host__ __device__ constexpr int fast_off(int k)
{
    const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    const int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    return dy[k] * FS_P + dx[k];
}

// segment test: >= 9 contiguous circle pixels brighter than v + t or darker than v - t
__device__ __forceinline__ bool fast_segment_test(const uint8_t *p, int threshold)
{
    // offsets relative to p in the staged tile (pitch FS_P)
    const int v = p[0];
    {
        // An arc of 9 of the 16 circle pixels contains one pixel of every antipodal pair: two pairs reject most
        // pixels after four loads (the same necessary condition cv::FAST tests first).
        const int e0 = v - p[fast_off(0)], e8 = v - p[fast_off(8)], e4 = v - p[fast_off(4)], e12 = v - p[fast_off(12)];
        const bool bright = (e0 < -threshold || e8 < -threshold) && (e4 < -threshold || e12 < -threshold);
        const bool dark = (e0 > threshold || e8 > threshold) && (e4 > threshold || e12 > threshold);
        if (!bright && !dark) return false;
    }
    unsigned mb = 0, md = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int d = v - p[fast_off(k)];
        mb |= (unsigned)(d < -threshold) << k;
        md |= (unsigned)(d > threshold) << k;
    }
    auto run9 = [](unsigned m) {
        m |= m << 16;
        unsigned a = m & (m >> 1);
        unsigned b = a & (a >> 2);
        unsigned c = b & (b >> 4);
        return (c & (m >> 8)) != 0u;
    };
    return run9(mb) || run9(md);
}

// cornerScore<16>: the largest threshold for which the pixel still passes the segment test
__device__ __forceinline__ int fast_corner_score(const uint8_t *p, int threshold)
{
    const int v = p[0];
    int d[25];
#pragma unroll
    for (int k = 0; k < 16; k++) d[k] = v - p[fast_off(k)];
#pragma unroll
    for (int k = 16; k < 25; k++) d[k] = d[k - 16];
    int a0 = threshold;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        int a = min(min(d[k + 1], d[k + 2]), d[k + 3]);
        if (a <= a0) continue;
        a = min(a, min(min(d[k + 4], d[k + 5]), min(d[k + 6], min(d[k + 7], d[k + 8]))));
        a0 = max(a0, min(a, d[k]));
        a0 = max(a0, min(a, d[k + 9]));
    }
    int b0 = -a0;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        int b = max(max(d[k + 1], d[k + 2]), max(d[k + 3], max(d[k + 4], d[k + 5])));
        if (b >= b0) continue;
        b = max(b, max(d[k + 6], max(d[k + 7], d[k + 8])));
        b0 = min(b0, max(b, d[k]));
        b0 = min(b0, max(b, d[k + 9]));
    }
    return -b0 - 1;
}

__global__ void __launch_bounds__(256)
fast_kernel(const uint8_t *__restrict__ img, int rows, int cols, int pitch, int threshold, int nonmax,
            Rec128 *__restrict__ out, int *__restrict__ count, int cap)
{
    __shared__ __align__(16) uint8_t s_px[FS_H][FS_P];
    __shared__ int s_sc[FC_H][FC_W + 1];
    __shared__ Rec128 s_rec[FT_H * FT_W];
    __shared__ unsigned short s_list[FC_H * FC_W];
    __shared__ int s_n, s_base, s_nl;
    if (threadIdx.x == 0) { s_n = 0; s_nl = 0; }   // the barriers below order it before the first use
    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * FT_W, ty0 = blockIdx.y * FT_H;
    if (tx0 >= 4 && tx0 + FT_W + 4 <= pitch && ty0 >= 4 && ty0 + FT_H + 4 <= rows) {
        // interior tile: the staged 72-byte rows start on a 4-byte boundary -> aligned 32-bit loads
        const uint8_t *g = img + (size_t)(ty0 - 4) * pitch + (tx0 - 4);
        for (int i = tid; i < FS_H * (FS_P / 4); i += 256) {
            const int r = i / (FS_P / 4), wc = i - r * (FS_P / 4);
            reinterpret_cast<uint32_t *>(&s_px[r][0])[wc] = __ldg(reinterpret_cast<const uint32_t *>(g + (size_t)r * pitch) + wc);
        }
    } else {
        for (int i = tid; i < FS_H * FS_W; i += 256) {
            int r = i / FS_W, c = i - r * FS_W;
            int gy = min(max(ty0 - 4 + r, 0), rows - 1), gx = min(max(tx0 - 4 + c, 0), cols - 1);
            s_px[r][c] = __ldg(img + (size_t)gy * pitch + gx);   // clamped reads are never used by valid pixels
        }
    }
    __syncthreads();
    // Segment test on the tile + 1 halo (pixels outside rows/cols 3..dim-4 are not corners); the pixels that pass are
    // LISTED, and the scores -- 10x the work of the test, needed by a few percent of the pixels -- are then computed by
    // the whole CTA from that list instead of by the odd lane of a diverged warp.
    for (int i = tid; i < FC_H * FC_W; i += 256) {
        int r = i / FC_W, c = i - r * FC_W;
        int y = ty0 - 1 + r, x = tx0 - 1 + c;
        s_sc[r][c] = 0;
        if (y >= 3 && y < rows - 3 && x >= 3 && x < cols - 3 && fast_segment_test(&s_px[r + 3][c + 3], threshold))
            s_list[atomicAdd(&s_nl, 1)] = (unsigned short)i;
    }
    __syncthreads();
    for (int k = tid; k < s_nl; k += 256) {
        const int i = s_list[k], r = i / FC_W, c = i - r * FC_W;
        s_sc[r][c] = fast_corner_score(&s_px[r + 3][c + 3], threshold);
    }
    __syncthreads();
    for (int i = tid; i < FT_H * FT_W; i += 256) {
        int r = i / FT_W, c = i - r * FT_W;
        int y = ty0 + r, x = tx0 + c;
        if (y >= rows || x >= cols) continue;
        int sc = s_sc[r + 1][c + 1];
        if (sc <= 0) continue;
        if (nonmax) {
            bool keep = sc > s_sc[r][c] && sc > s_sc[r][c + 1] && sc > s_sc[r][c + 2] && sc > s_sc[r + 1][c] &&
                        sc > s_sc[r + 1][c + 2] && sc > s_sc[r + 2][c] && sc > s_sc[r + 2][c + 1] && sc > s_sc[r + 2][c + 2];
            if (!keep) continue;
        }
        // gathered per tile: the global counter takes one atomic per CTA (same-address atomics serialise in L2)
        const int slot = atomicAdd(&s_n, 1);
        s_rec[slot] = Rec128{~(unsigned long long)(y * cols + x), (unsigned long long)(nonmax ? sc : 0)};
    }
    __syncthreads();
    const int nloc = s_n;
    if (nloc == 0) return;
    if (tid == 0) s_base = atomicAdd(count, nloc);
    __syncthreads();
    const int base = s_base;
    for (int i = tid; i < nloc; i += 256)
        if (base + i < cap) out[base + i] = s_rec[i];
}

__global__ void __launch_bounds__(256)
fast_emit_kernel(const Rec128 *__restrict__ recs, int n, int cols, int *__restrict__ col, int *__restrict__ row,
                 float *__restrict__ score)
{
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    int idx = (int)(~recs[i].hi);
    row[i] = idx / cols; col[i] = idx - (idx / cols) * cols;
    score[i] = (float)(int)recs[i].lo;
}

}  // namespace

extern "C" {

PMV_API int pmv_fast(pmv_ctx *ctx, const uint8_t *img, int rows, int cols, int step, int threshold, int nonmax,
                     int max_feats, int *col, int *row, float *score, int *n_out, int *n_total)
{
    if (!ctx) return PMV_ERR_INVALID;
    if (!img || rows <= 0 || cols <= 0 || step < cols || !n_out || max_feats < 0 || threshold < 0 || threshold > 255)
        return ctx->fail(PMV_ERR_INVALID, "pmv_fast: bad argument");
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    const size_t npx = (size_t)rows * cols;
    const int pitch = align_up(cols, 128);
    const int cap = (int)npx;
    cudaError_t e = ctx->img[0].reserve((size_t)pitch * rows);
    if (e == cudaSuccess) e = ctx->scratch[1].reserve(256);
    if (e == cudaSuccess) e = ctx->scratch[2].reserve((size_t)sort_capacity(cap) * sizeof(Rec128));
    if (e == cudaSuccess) e = ctx->pin[0].reserve(64);
    if (e != cudaSuccess) return ctx->fail(PMV_ERR_NOMEM, "fast workspace", e);
    PMV_CUDA_TRY(ctx, cudaMemcpy2DAsync(ctx->img[0].p, pitch, img, step, cols, rows, cudaMemcpyHostToDevice, s));
    int *d_count = ctx->scratch[1].as<int>();
    Rec128 *d_rec = ctx->scratch[2].as<Rec128>();
    int *h_misc = ctx->pin[0].as<int>();
    PMV_CUDA_TRY(ctx, cudaMemsetAsync(d_count, 0, 16, s));
    {
        ProfScope ps(ctx, PMV_PHASE_FAST, s);
        dim3 grid((cols + FT_W - 1) / FT_W, (rows + FT_H - 1) / FT_H);
        fast_kernel<<<grid, 256, 0, s>>>(ctx->img[0].as<uint8_t>(), rows, cols, pitch, threshold, nonmax, d_rec, d_count, cap);
        PMV_LAUNCH_CHECK(ctx, "fast_kernel");
    }
    PMV_CUDA_TRY(ctx, cudaMemcpyAsync(h_misc, d_count, 4, cudaMemcpyDeviceToHost, s));
    PMV_CUDA_TRY(ctx, cudaStreamSynchronize(s));
    const int total = h_misc[0];
    if (n_total) *n_total = total;
    const int n = total < max_feats ? total : max_feats;
    if (n > 0) {
        ProfScope ps(ctx, PMV_PHASE_SELECT, s);
        int rc = PMV_OK;
        // raster order = descending ~index.  Large lists: buckets of 2^shift consecutive pixel indices (at most that many
        // keypoints each, so the quadratic last pass is bounded) -- five small launches instead of the bitonic network's
        // 36 at 2^18 records (sort.cuh)
        int shift = 9;
        while (((npx - 1) >> shift) + 2 > (size_t)BS_BINS) shift++;    // bucket = (~0 >> shift) - (~index >> shift) <= (index >> shift) + 1
        if (total >= 4 * SORT_CHUNK && shift <= 11) {
            e = ctx->scratch[3].reserve((size_t)total * sizeof(Rec128));
            if (e == cudaSuccess) e = ctx->scratch[4].reserve(3 * (size_t)BS_BINS * sizeof(int) + 16);
            if (e != cudaSuccess) return ctx->fail(PMV_ERR_NOMEM, "fast sort workspace", e);
            Rec128 *d_tmp = ctx->scratch[3].as<Rec128>();
            int *d_hist = ctx->scratch[4].as<int>(), *d_start = d_hist + BS_BINS, *d_cursor = d_start + BS_BINS, *d_maxc = d_cursor + BS_BINS;
            PMV_CUDA_TRY(ctx, cudaMemsetAsync(d_hist, 0, 3 * (size_t)BS_BINS * sizeof(int) + 16, s));
            bs_hist_kernel<<<(total + 255) / 256, 256, 0, s>>>(d_rec, total, 0xffffffffu, shift, d_hist);
            PMV_LAUNCH_CHECK(ctx, "bs_hist_kernel");
            if (ctx->attr_first(PMV_ATTR_BS_SCAN_FAST)) cudaFuncSetAttribute(bs_scan_kernel,   // this translation unit's copy of the kernel
                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BS_SCAN_SMEM);
            bs_scan_kernel<<<1, 1024, BS_SCAN_SMEM, s>>>(d_hist, d_start, d_maxc);
            PMV_LAUNCH_CHECK(ctx, "bs_scan_kernel");
            bs_scatter_kernel<<<(total + 255) / 256, 256, 0, s>>>(d_rec, total, nullptr, d_start, d_cursor, d_tmp, 0xffffffffu, shift);
            PMV_LAUNCH_CHECK(ctx, "bs_scatter_kernel");
            bs_rank_kernel<<<(total + 255) / 256, 256, 0, s>>>(d_tmp, total, nullptr, d_start, d_hist, d_rec, 0xffffffffu, shift);
            PMV_LAUNCH_CHECK(ctx, "bs_rank_kernel");
        } else {
            rc = sort_desc_128(ctx, d_rec, total, s);
        }
        if (rc) return rc;
        e = ctx->scratch[5].reserve((size_t)n * 12 + 16);
        if (e != cudaSuccess) return ctx->fail(PMV_ERR_NOMEM, "fast output", e);
        int *d_col = ctx->scratch[5].as<int>(), *d_row = d_col + n;
        float *d_sc = reinterpret_cast<float *>(d_row + n);
        fast_emit_kernel<<<(n + 255) / 256, 256, 0, s>>>(d_rec, n, cols, d_col, d_row, d_sc);
        PMV_LAUNCH_CHECK(ctx, "fast_emit_kernel");
        if (col) PMV_CUDA_TRY(ctx, cudaMemcpyAsync(col, d_col, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
        if (row) PMV_CUDA_TRY(ctx, cudaMemcpyAsync(row, d_row, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
        if (score) PMV_CUDA_TRY(ctx, cudaMemcpyAsync(score, d_sc, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
        PMV_CUDA_TRY(ctx, cudaStreamSynchronize(s));
    }
    *n_out = n;
    return PMV_OK;
}

}  // extern "C"
