// pyramid.cu -- K1 Gaussian pyramid (cv::pyrDown chain) and K2 Scharr derivative image.
//
// Replaces the buildOpticalFlowPyramid / calcSharrDeriv work hidden inside
// cv::calcOpticalFlowPyrLK at reference OpenCVLucasKanadeFM.cpp:15 (SURVEY Appx A.1, A.2).
// Integer arithmetic, bit-exact: separable [1 4 6 4 1] taps centred on even source pixels,
// BORDER_REFLECT_101, (sum + 128) >> 8.
//
// HBM-bound streaming kernel: one CTA produces a 128x16 output tile from a 259x35 input tile
// staged in shared memory with 16-byte vector loads (rows of context-owned levels are pitched
// to 128 B), horizontal pass into a uint16 tile, vertical pass + packed 4-byte stores.
// Algorithmic bytes per image: W*H read + sum_l W_l*H_l written.
#include "common.cuh"

namespace {

constexpr int PT_W = 128;               // output tile width
constexpr int PT_H = 16;                // output tile height
constexpr int PIN_H = 2 * PT_H + 3;     // 35 input rows
constexpr int PIN_W = 2 * PT_W + 32;    // 288 staged input bytes per row (16 B aligned superset)
constexpr int PIN_X0 = 16;              // staged column c <-> global column 2*tx0 - 16 + c

__global__ void __launch_bounds__(256)
pyr_down_kernel(const uint8_t *__restrict__ src, int srows, int scols, int spitch, size_t sstride,
                uint8_t *__restrict__ dst, int drows, int dcols, int dpitch, size_t dstride,
                int vec_ok)
{
    __shared__ __align__(16) uint8_t s_in[PIN_H][PIN_W];
    __shared__ __align__(16) uint16_t s_h[PIN_H][PT_W];

    const int b = blockIdx.z;
    src += (size_t)b * sstride;
    dst += (size_t)b * dstride;
    const int tx0 = blockIdx.x * PT_W, ty0 = blockIdx.y * PT_H;
    const int gx0 = 2 * tx0 - PIN_X0;  // global column of staged column 0
    const int gy0 = 2 * ty0 - 2;       // global row of staged row 0
    const int tid = threadIdx.x;

    // ---- stage input tile ------------------------------------------------------------
    if (vec_ok) {
        // 18 x uint4 per row; chunks fully inside [0, spitch) come straight from memory
        for (int i = tid; i < PIN_H * (PIN_W / 16); i += 256) {
            int r = i / (PIN_W / 16), ch = i % (PIN_W / 16);
            int gy = reflect101(gy0 + r, srows);
            int gx = gx0 + ch * 16;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (gx >= 0 && gx + 16 <= spitch)
                v = __ldg(reinterpret_cast<const uint4 *>(src + (size_t)gy * spitch + gx));
            *reinterpret_cast<uint4 *>(&s_in[r][ch * 16]) = v;
        }
        __syncthreads();
        // fix-up: columns outside [0, scols) that the taps can touch (reflect-101)
        const bool edge = (gx0 + PIN_X0 - 2 < 0) || (gx0 + PIN_X0 + 2 * PT_W + 2 > scols);
        if (edge) {
            for (int i = tid; i < PIN_H * (2 * PT_W + 3); i += 256) {
                int r = i / (2 * PT_W + 3), c = PIN_X0 - 2 + i % (2 * PT_W + 3);
                int gx = gx0 + c;
                if (gx < 0 || gx >= scols) {
                    int gy = reflect101(gy0 + r, srows);
                    s_in[r][c] = src[(size_t)gy * spitch + reflect101(gx, scols)];
                }
            }
        }
    } else {
        for (int i = tid; i < PIN_H * (2 * PT_W + 3); i += 256) {
            int r = i / (2 * PT_W + 3), c = PIN_X0 - 2 + i % (2 * PT_W + 3);
            int gy = reflect101(gy0 + r, srows);
            int gx = reflect101(gx0 + c, scols);
            s_in[r][c] = src[(size_t)gy * spitch + gx];
        }
    }
    __syncthreads();

    // ---- horizontal pass: s_h[r][x] = taps over s_in[r][2x-2 .. 2x+2] -------------------
    for (int i = tid; i < PIN_H * PT_W; i += 256) {
        int r = i / PT_W, x = i % PT_W;
        const uint8_t *p = &s_in[r][PIN_X0 + 2 * x - 2];
        s_h[r][x] = (uint16_t)(p[0] + p[4] + 4 * (p[1] + p[3]) + 6 * p[2]);
    }
    __syncthreads();

    // ---- vertical pass + store: a warp owns a row, a lane 4 consecutive outputs ---------
    const int lane = tid & 31, warp = tid >> 5;
    for (int yy = warp; yy < PT_H; yy += 8) {
        int oy = ty0 + yy;
        if (oy >= drows) break;
        uint32_t packed = 0;
        int ox = tx0 + lane * 4;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int x = lane * 4 + k;
            int v = s_h[2 * yy][x] + s_h[2 * yy + 4][x] + 4 * (s_h[2 * yy + 1][x] + s_h[2 * yy + 3][x]) +
                    6 * s_h[2 * yy + 2][x];
            packed |= (uint32_t)((v + 128) >> 8) << (8 * k);
        }
        uint8_t *o = dst + (size_t)oy * dpitch + ox;
        if (ox + 4 <= dpitch && (dpitch & 3) == 0) {
            // pitch padding absorbs the partial word at the right edge
            if (ox < dcols) *reinterpret_cast<uint32_t *>(o) = packed;
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (ox + k < dcols) o[k] = (uint8_t)(packed >> (8 * k));
        }
    }
}

// K2: int16 x2 Scharr derivative, reflect-101 (stage-by-stage parity entry point; the LK
// kernel computes the same values on the fly from its staged patch).
__global__ void __launch_bounds__(256)
scharr_kernel(const uint8_t *__restrict__ src, int rows, int cols, int pitch, short2 *__restrict__ dst)
{
    int x = blockIdx.x * 32 + (threadIdx.x & 31);
    int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= cols || y >= rows) return;
    int y0 = reflect101(y - 1, rows), y2 = reflect101(y + 1, rows);
    int xm = reflect101(x - 1, cols), xp = reflect101(x + 1, cols);
    const uint8_t *r0 = src + (size_t)y0 * pitch, *r1 = src + (size_t)y * pitch, *r2 = src + (size_t)y2 * pitch;
    int t0m = 3 * (r0[xm] + r2[xm]) + 10 * r1[xm];
    int t0p = 3 * (r0[xp] + r2[xp]) + 10 * r1[xp];
    int t1m = r2[xm] - r0[xm], t1c = r2[x] - r0[x], t1p = r2[xp] - r0[xp];
    dst[(size_t)y * cols + x] = make_short2((short)(t0p - t0m), (short)(3 * (t1m + t1p) + 10 * t1c));
}

}  // namespace

// ------------------------------------------------------------------ internal planning ---
// Lay out reduced levels 1..L of a batch in ctx->pyr[which]; level 0 aliases d_lvl0.
int pmv_internal_pyr_plan(pmv_ctx *ctx, int which, int batch, int rows, int cols,
                          const uint8_t *d_lvl0, int pitch0, size_t stride0,
                          int win_w, int win_h, int max_level, PyrSet *out)
{
    int L = pmv_pyr_levels(rows, cols, win_w, win_h, max_level);
    if (L < 0 || L >= PMV_MAX_PYR_LEVELS) return ctx->fail(PMV_ERR_UNSUPPORTED, "max_level too large");
    out->top = L;
    out->lv[0] = PyrLevel{d_lvl0, rows, cols, pitch0, stride0};
    size_t total = 0;
    size_t off[PMV_MAX_PYR_LEVELS] = {0};
    int r = rows, c = cols;
    for (int l = 1; l <= L; l++) {
        r = (r + 1) / 2;
        c = (c + 1) / 2;
        int pitch = align_up(c, 128);
        size_t stride = (size_t)pitch * r;
        off[l] = total;
        total += stride * batch;
        out->lv[l] = PyrLevel{nullptr, r, c, pitch, stride};
    }
    if (total) {
        cudaError_t e = ctx->pyr[which].reserve(total);
        if (e != cudaSuccess) return ctx->fail(PMV_ERR_NOMEM, "pyramid workspace", e);
    }
    for (int l = 1; l <= L; l++) out->lv[l].ptr = ctx->pyr[which].as<uint8_t>() + off[l];
    return PMV_OK;
}

int pmv_internal_pyr_run(pmv_ctx *ctx, const PyrSet &set, int batch, cudaStream_t s)
{
    for (int l = 1; l <= set.top; l++) {
        const PyrLevel &a = set.lv[l - 1], &d = set.lv[l];
        int vec_ok = (a.pitch % 16 == 0) && (((uintptr_t)a.ptr) % 16 == 0) && (a.img_stride % 16 == 0);
        dim3 grid((d.cols + PT_W - 1) / PT_W, (d.rows + PT_H - 1) / PT_H, batch);
        pyr_down_kernel<<<grid, 256, 0, s>>>(a.ptr, a.rows, a.cols, a.pitch, a.img_stride,
                                             const_cast<uint8_t *>(d.ptr), d.rows, d.cols, d.pitch,
                                             d.img_stride, vec_ok);
        PMV_LAUNCH_CHECK(ctx, "pyr_down_kernel");
    }
    return PMV_OK;
}

// ------------------------------------------------------------------ C ABI ---------------
extern "C" {

PMV_API int pmv_pyr_levels(int rows, int cols, int win_w, int win_h, int max_level)
{
    if (rows <= 0 || cols <= 0 || max_level < 0) return -1;
    int w = cols, h = rows, level = 0;
    for (level = 0; level < max_level; level++) {
        int nw = (w + 1) / 2, nh = (h + 1) / 2;
        if (nw <= win_w || nh <= win_h) break;
        w = nw;
        h = nh;
    }
    return level;
}

PMV_API int pmv_pyramid_build(pmv_ctx *ctx, const uint8_t *img, int rows, int cols, int step,
                              int win_w, int win_h, int max_level,
                              uint8_t *out_packed, size_t out_capacity, int *out_levels)
{
    if (!ctx) return PMV_ERR_INVALID;
    if (!img || rows <= 0 || cols <= 0 || step < cols || max_level < 0 || !out_levels)
        return ctx->fail(PMV_ERR_INVALID, "pmv_pyramid_build: bad argument");
    cudaSetDevice(ctx->device);
    int pitch0 = align_up(cols, 128);
    cudaError_t e = ctx->img[0].reserve((size_t)pitch0 * rows);
    if (e != cudaSuccess) return ctx->fail(PMV_ERR_NOMEM, "image upload buffer", e);
    PMV_CUDA_TRY(ctx, cudaMemcpy2DAsync(ctx->img[0].p, pitch0, img, step, cols, rows,
                                        cudaMemcpyHostToDevice, ctx->stream));
    PyrSet set;
    int rc = pmv_internal_pyr_plan(ctx, 0, 1, rows, cols, ctx->img[0].as<uint8_t>(), pitch0,
                                   (size_t)pitch0 * rows, win_w, win_h, max_level, &set);
    if (rc) return rc;
    rc = pmv_internal_pyr_run(ctx, set, 1, ctx->stream);
    if (rc) return rc;
    size_t need = 0;
    for (int l = 1; l <= set.top; l++) need += (size_t)set.lv[l].rows * set.lv[l].cols;
    if (need > out_capacity || (need && !out_packed))
        return ctx->fail(PMV_ERR_INVALID, "pmv_pyramid_build: output buffer too small");
    size_t o = 0;
    for (int l = 1; l <= set.top; l++) {
        const PyrLevel &d = set.lv[l];
        PMV_CUDA_TRY(ctx, cudaMemcpy2DAsync(out_packed + o, d.cols, d.ptr, d.pitch, d.cols, d.rows,
                                            cudaMemcpyDeviceToHost, ctx->stream));
        o += (size_t)d.rows * d.cols;
    }
    PMV_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *out_levels = set.top;
    return PMV_OK;
}

PMV_API int pmv_scharr(pmv_ctx *ctx, const uint8_t *img, int rows, int cols, int step, int16_t *out)
{
    if (!ctx) return PMV_ERR_INVALID;
    if (!img || !out || rows <= 0 || cols <= 0 || step < cols)
        return ctx->fail(PMV_ERR_INVALID, "pmv_scharr: bad argument");
    cudaSetDevice(ctx->device);
    int pitch0 = align_up(cols, 128);
    cudaError_t e = ctx->img[0].reserve((size_t)pitch0 * rows);
    if (e == cudaSuccess) e = ctx->scratch[0].reserve((size_t)rows * cols * 4);
    if (e != cudaSuccess) return ctx->fail(PMV_ERR_NOMEM, "scharr buffers", e);
    PMV_CUDA_TRY(ctx, cudaMemcpy2DAsync(ctx->img[0].p, pitch0, img, step, cols, rows,
                                        cudaMemcpyHostToDevice, ctx->stream));
    dim3 grid((cols + 31) / 32, (rows + 7) / 8);
    scharr_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->img[0].as<uint8_t>(), rows, cols, pitch0,
                                                 ctx->scratch[0].as<short2>());
    PMV_LAUNCH_CHECK(ctx, "scharr_kernel");
    PMV_CUDA_TRY(ctx, cudaMemcpyAsync(out, ctx->scratch[0].p, (size_t)rows * cols * 4,
                                      cudaMemcpyDeviceToHost, ctx->stream));
    PMV_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return PMV_OK;
}

}  // extern "C"
