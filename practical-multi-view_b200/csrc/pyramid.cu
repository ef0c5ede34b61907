// pyramid.cu -- K1 Gaussian pyramid (cv::pyrDown chain) and K2 Scharr derivative image.
//
// Replaces the buildOpticalFlowPyramid / calcSharrDeriv work hidden inside
// cv::calcOpticalFlowPyrLK at reference OpenCVLucasKanadeFM.cpp:15 (SURVEY Appx A.1, A.2).
// Integer arithmetic, bit-exact: separable [1 4 6 4 1] taps centred on even source pixels,
// BORDER_REFLECT_101, (sum + 128) >> 8.
//
// Storage: like OpenCV's optical-flow pyramid every level (level 0 included) is kept with a
// reflect-101 border, in rows pitched to 128 B.  The border makes both the pyrDown taps and all
// LK patch reads plain, aligned loads (no coordinate reflection in the hot loops).
//
// Kernels (all HBM-bound streaming kernels, DESIGN.md has the byte counts):
//   import_kernel      caller image (any pitch) -> interior of the bordered level 0
//   border_fill_kernel reflect-101 border of one level, all images of the batch
//   pyr_down_kernel    register-resident: a warp = 60 x 8 outputs, packed 16-bit SIMD-in-register
//                      arithmetic, neighbour sums by warp shuffles, no shared memory
#include "common.cuh"

namespace {

constexpr int PW_OUT = 60;   // outputs per warp-row: lanes 1..30 produce two each (lanes 0 / 31 are halo)
constexpr int PW_R = 8;      // output rows per warp

// One aligned 32-bit word of a level row; words touching columns outside [0, cols) are assembled
// byte-wise with reflect-101 (edge lanes only).
__device__ __forceinline__ uint32_t pyr_load_word(const uint8_t *__restrict__ row, int col0, int cols)
{
    if (col0 >= 0 && col0 + 4 <= cols) return __ldg(reinterpret_cast<const uint32_t *>(row + col0));
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int c = col0 + k;
        c = c < -2 ? 0 : (c > cols + 1 ? cols - 1 : reflect101(c, cols));   // beyond the taps: any in-range byte
        v |= (uint32_t)__ldg(row + c) << (8 * k);
    }
    return v;
}

// K1 pyrDown, register-resident: a warp owns 60 output columns x PW_R output rows.  Lane L holds the
// 4-byte input word at columns 2*x0 - 4 + 4L; the vertical [1 4 6 4 1] pass runs on two packed 16-bit
// lanes per register (even / odd input columns, sums <= 4080), the horizontal pass on packed output
// pairs (sums <= 65280) with the neighbouring lanes' partial sums fetched by shuffles.  No shared
// memory, no barriers; every input word is read once per warp (+ 2 halo lanes, + 3 halo rows per 16).
__global__ void __launch_bounds__(256)
pyr_down_kernel(const uint8_t *__restrict__ src, int srows, int scols, int spitch, size_t sstride,
                uint8_t *__restrict__ dst, int drows, int dcols, int dpitch, size_t dstride)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x0 = blockIdx.x * PW_OUT;                         // first output column of the warp
    const int y0 = (blockIdx.y * 8 + warp) * PW_R;              // first output row of the warp
    if (y0 >= drows) return;
    src += (size_t)blockIdx.z * sstride;
    dst += (size_t)blockIdx.z * dstride;
    const int col0 = 2 * x0 - 4 + 4 * lane;                     // input column of byte 0 of this lane's word
    const uint32_t M = 0x00ff00ffu;

    // all 2*PW_R + 3 input rows of the warp are requested up front (independent loads: enough bytes in
    // flight to cover HBM latency), rows past the image fold back through reflect-101 and are harmless
    uint32_t win[2 * PW_R + 3];
#pragma unroll
    for (int k = 0; k < 2 * PW_R + 3; k++) {
        const uint8_t *row = src + (size_t)reflect101(2 * y0 - 2 + k, srows) * spitch;
        win[k] = pyr_load_word(row, col0, scols);
    }
#pragma unroll
    for (int r = 0; r < PW_R; r++) {
        const int y = y0 + r;
        if (y >= drows) break;                                  // warp-uniform
        uint32_t e[5], o[5];                                    // rows 2y-2 .. 2y+2, even / odd columns
#pragma unroll
        for (int k = 0; k < 5; k++) { e[k] = win[2 * r + k] & M; o[k] = (win[2 * r + k] >> 8) & M; }
        // vertical pass (two input columns per register)
        const uint32_t VE = e[0] + e[4] + ((e[1] + e[3]) << 2) + e[2] * 6u;   // (V0, V2)
        const uint32_t VO = o[0] + o[4] + ((o[1] + o[3]) << 2) + o[2] * 6u;   // (V1, V3)
        const uint32_t VEl = __shfl_up_sync(0xffffffffu, VE, 1), VOl = __shfl_up_sync(0xffffffffu, VO, 1);
        const uint32_t VEr = __shfl_down_sync(0xffffffffu, VE, 1);
        // horizontal pass on the output pair (x, x+1), x = x0 + 2 (lane - 1)
        const uint32_t A = __byte_perm(VEl, VE, 0x5432);        // (V[-2], V0)
        const uint32_t B = __byte_perm(VOl, VO, 0x5432);        // (V[-1], V1)
        const uint32_t F = __byte_perm(VE, VEr, 0x5432);        // (V2, V4)
        const uint32_t res = A + F + ((B + VO) << 2) + VE * 6u + 0x00800080u;
        const uint32_t two = __byte_perm((res >> 8) & M, 0u, 0x4420);   // out(x) | out(x+1) << 8
        const uint32_t nb = __shfl_down_sync(0xffffffffu, two, 1);
        if ((lane & 1) && lane < 31) {
            const int ox = x0 + 2 * (lane - 1);
            if (ox < dcols) {
                const uint32_t four = two | (nb << 16);
                // row pitch padding absorbs a partial word at the right edge (border fill follows)
                *reinterpret_cast<uint32_t *>(dst + (size_t)y * dpitch + ox) = four;
            }
        }
    }
}

// Reflect-101 border of ALL levels of a batch in one launch (grid.y = image, grid.z = level):
// bands top / bottom (full bordered width) and left / right.
struct BorderArgs {
    uint8_t *ptr[PMV_MAX_PYR_LEVELS];
    int rows[PMV_MAX_PYR_LEVELS], cols[PMV_MAX_PYR_LEVELS], pitch[PMV_MAX_PYR_LEVELS];
    size_t stride[PMV_MAX_PYR_LEVELS];
    int by, bxl, bxr;
};

// Work item = one aligned 16-byte chunk of a destination row.  Border rows (above / below the image) copy
// whole chunks from their mirror row with one uint4 load where the chunk lies inside [0, cols); everything
// else (left / right margins, row ends) is assembled byte-wise with reflect-101.
__global__ void __launch_bounds__(256) border_fill_kernel(const BorderArgs A)
{
    const int l = blockIdx.z;
    const int rows = A.rows[l], cols = A.cols[l], pitch = A.pitch[l];
    uint8_t *base = A.ptr[l] + (size_t)blockIdx.y * A.stride[l];
    const int by = A.by, bxl = A.bxl, bxr = A.bxr;
    const int row_chunks = (bxl + cols + bxr + 15) >> 4;          // chunks of a full bordered row
    const int right0 = cols & ~15;                                // first chunk (interior x) touching the right margin
    const int side_chunks = (bxl >> 4) + ((cols + bxr + 15 - right0) >> 4);
    const int n_tb = 2 * by * row_chunks, n_lr = rows * side_chunks;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n_tb + n_lr; i += gridDim.x * 256) {
        int y, x0;
        if (i < n_tb) {
            const int r = i / row_chunks;
            x0 = (i - r * row_chunks) * 16 - bxl;
            y = r < by ? r - by : rows + (r - by);
        } else {
            const int j = i - n_tb;
            y = j / side_chunks;
            const int c = j - y * side_chunks;
            x0 = c < (bxl >> 4) ? c * 16 - bxl : right0 + (c - (bxl >> 4)) * 16;
        }
        const uint8_t *srow = base + (ptrdiff_t)reflect101(y, rows) * pitch;
        uint8_t *drow = base + (ptrdiff_t)y * pitch;
        if (x0 >= 0 && x0 + 16 <= cols) {
            if (y < 0 || y >= rows) *reinterpret_cast<uint4 *>(drow + x0) = *reinterpret_cast<const uint4 *>(srow + x0);
        } else {
            uint32_t wv[4];
#pragma unroll
            for (int k4 = 0; k4 < 4; k4++) {
                uint32_t v = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int x = x0 + 4 * k4 + k;
                    v |= (uint32_t)srow[reflect101(x, cols)] << (8 * k);
                }
                wv[k4] = v;
            }
            if (y >= 0 && y < rows && (x0 + 16 <= 0 || (x0 >= cols && x0 + 16 <= cols + bxr))) {
                *reinterpret_cast<uint4 *>(drow + x0) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
            } else if (y >= 0 && y < rows) {
                // interior row: keep the image bytes of a chunk that straddles the edge
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    const int x = x0 + k;
                    if ((x < 0 || x >= cols) && x < cols + bxr) drow[x] = (uint8_t)(wv[k >> 2] >> (8 * (k & 3)));
                }
            } else if (x0 + 16 <= cols + bxr) {
                *reinterpret_cast<uint4 *>(drow + x0) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
            } else {
#pragma unroll
                for (int k = 0; k < 16; k++)
                    if (x0 + k < cols + bxr) drow[x0 + k] = (uint8_t)(wv[k >> 2] >> (8 * (k & 3)));
            }
        }
    }
}

// Caller image (device memory, any pitch / alignment) -> interior of bordered level 0.
// 16 bytes per thread: one uint4 load when the source is 16 B aligned and pitched, else byte gathers.
__global__ void __launch_bounds__(256)
import_kernel(const uint8_t *__restrict__ src, int spitch, size_t sstride, int rows, int cols,
              uint8_t *__restrict__ dst, int dpitch, size_t dstride, int src_aligned16)
{
    const int x16 = (blockIdx.x * 32 + (threadIdx.x & 31)) * 16;
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (y >= rows || x16 >= cols) return;
    const uint8_t *sp = src + (size_t)blockIdx.z * sstride + (size_t)y * spitch + x16;
    uint4 v;
    if (src_aligned16 && x16 + 16 <= spitch) {
        v = __ldg(reinterpret_cast<const uint4 *>(sp));
    } else {
        uint32_t wv[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 16; k++)
            if (x16 + k < cols) wv[k >> 2] |= (uint32_t)__ldg(sp + k) << (8 * (k & 3));
        v = make_uint4(wv[0], wv[1], wv[2], wv[3]);
    }
    // the right border of the destination absorbs the tail of the last chunk (filled afterwards)
    *reinterpret_cast<uint4 *>(dst + (size_t)blockIdx.z * dstride + (size_t)y * dpitch + x16) = v;
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

// K2 for the LK tracker: derivative of every pixel of a bordered level, 4 pixels per thread (three aligned
// words per input row, one 16 B store).  Reads run 1 px into the reflect-101 border of the level, so the
// values at the image edge are OpenCV's; pixels beyond the last column are written as zero (they belong
// to the zero border of the derivative plane).
__global__ void __launch_bounds__(256)
scharr_level_kernel(const uint8_t *__restrict__ img, int rows, int cols, int pitch, size_t istride,
                    int *__restrict__ der, int dpitch, size_t dstride)
{
    const int x0 = (blockIdx.x * 64 + (threadIdx.x & 63)) * 4;
    const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x0 >= cols || y >= rows) return;
    const uint8_t *base = img + (size_t)blockIdx.z * istride + (size_t)(y - 1) * pitch + x0 - 4;
    uint32_t wv[3][3];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int k = 0; k < 3; k++) wv[r][k] = *reinterpret_cast<const uint32_t *>(base + (size_t)r * pitch + 4 * k);
    int t0[6], t1[6];   // columns x0 - 1 .. x0 + 4
#pragma unroll
    for (int c = 0; c < 6; c++) {
        const int bi = c + 3;   // byte index in the 12-byte row
        const int a0 = (wv[0][bi >> 2] >> (8 * (bi & 3))) & 255, a1 = (wv[1][bi >> 2] >> (8 * (bi & 3))) & 255,
                  a2 = (wv[2][bi >> 2] >> (8 * (bi & 3))) & 255;
        t0[c] = 3 * (a0 + a2) + 10 * a1;
        t1[c] = a2 - a0;
    }
    int out[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int gx = t0[k + 2] - t0[k], gy = 3 * (t1[k] + t1[k + 2]) + 10 * t1[k + 1];
        out[k] = (x0 + k < cols) ? ((gx & 0xffff) | (int)((unsigned)gy << 16)) : 0;
    }
    *reinterpret_cast<int4 *>(der + (size_t)blockIdx.z * dstride + (size_t)y * dpitch + x0) = make_int4(out[0], out[1], out[2], out[3]);
}

}  // namespace

int pmv_internal_deriv_plan(pmv_ctx *ctx, const PyrSet &set, int batch, DerivSet *out, cudaStream_t s)
{
    size_t total = 0, off[PMV_MAX_PYR_LEVELS];
    unsigned long long sig = 1469598103934665603ull;
    auto mix = [&](unsigned long long v) { sig = (sig ^ v) * 1099511628211ull; };
    mix((unsigned long long)batch); mix((unsigned long long)set.top);
    for (int l = 0; l <= set.top; l++) {
        const PyrLevel &a = set.lv[l];
        const int bd = a.border, bl = align_up(bd, 4);
        const int pitch = align_up(bl + a.cols + bd + 4, 32);   // +4: the last 16 B store may run past the last column
        const size_t stride = (size_t)pitch * (a.rows + 2 * bd);
        off[l] = total + (size_t)bd * pitch + bl;
        total += stride * batch;
        out->lv[l] = DerivLevel{nullptr, pitch, stride};
        mix((unsigned long long)a.rows); mix((unsigned long long)a.cols); mix((unsigned long long)bd);
    }
    const void *before = ctx->deriv.p;
    cudaError_t e = ctx->deriv.reserve(total * 4 + 256);
    if (e != cudaSuccess) return ctx->fail(PMV_ERR_NOMEM, "derivative workspace", e);
    if (ctx->deriv.p != before || ctx->deriv_sig != sig) {   // new geometry: (re)write the zero borders
        PMV_CUDA_TRY(ctx, cudaMemsetAsync(ctx->deriv.p, 0, total * 4, s));
        ctx->deriv_sig = sig;
    }
    for (int l = 0; l <= set.top; l++) out->lv[l].ptr = ctx->deriv.as<int>() + off[l];
    return PMV_OK;
}

int pmv_internal_deriv_run(pmv_ctx *ctx, const PyrSet &set, const DerivSet &d, int batch, cudaStream_t s)
{
    for (int l = 0; l <= set.top; l++) {
        const PyrLevel &a = set.lv[l];
        dim3 grid((a.cols + 255) / 256, (a.rows + 3) / 4, batch);
        scharr_level_kernel<<<grid, 256, 0, s>>>(a.ptr, a.rows, a.cols, a.pitch, a.img_stride,
                                                 const_cast<int *>(d.lv[l].ptr), d.lv[l].pitch, d.lv[l].img_stride);
        PMV_LAUNCH_CHECK(ctx, "scharr_level_kernel");
    }
    return PMV_OK;
}

// ------------------------------------------------------------------ internal planning ---
int pmv_internal_pyr_plan(pmv_ctx *ctx, int which, int batch, int rows, int cols, int border,
                          int win_w, int win_h, int max_level, PyrSet *out)
{
    int L = pmv_pyr_levels(rows, cols, win_w, win_h, max_level);
    if (L < 0 || L >= PMV_MAX_PYR_LEVELS) return ctx->fail(PMV_ERR_UNSUPPORTED, "max_level too large");
    if (border < 2) border = 2;
    out->top = L;
    const int bxl = align_up(border + 4, 16);  // left border: keeps the interior 16 B aligned
    size_t total = 0;
    size_t off[PMV_MAX_PYR_LEVELS];
    int r = rows, c = cols;
    for (int l = 0; l <= L; l++) {
        if (l > 0) { r = (r + 1) / 2; c = (c + 1) / 2; }
        int pitch = align_up(bxl + c + border + 16, 128);   // +16: 16-byte stores may run past the right border
        size_t stride = (size_t)pitch * (r + 2 * border);
        off[l] = total + (size_t)border * pitch + bxl;  // interior origin of image 0
        total += stride * batch;
        out->lv[l] = PyrLevel{nullptr, r, c, pitch, stride, border, bxl};
    }
    cudaError_t e = ctx->pyr[which].reserve(total + 256);
    if (e != cudaSuccess) return ctx->fail(PMV_ERR_NOMEM, "pyramid workspace", e);
    for (int l = 0; l <= L; l++) out->lv[l].ptr = ctx->pyr[which].as<uint8_t>() + off[l];
    return PMV_OK;
}

static int fill_borders(pmv_ctx *ctx, const PyrSet &set, int batch, cudaStream_t s)
{
    BorderArgs A;
    memset(&A, 0, sizeof A);
    int nmax = 0;
    for (int l = 0; l <= set.top; l++) {
        const PyrLevel &d = set.lv[l];
        A.ptr[l] = const_cast<uint8_t *>(d.ptr); A.rows[l] = d.rows; A.cols[l] = d.cols; A.pitch[l] = d.pitch;
        A.stride[l] = d.img_stride;
        int n = 2 * d.border * (d.bxl + d.cols + d.border) + d.rows * (d.bxl + d.border);
        nmax = n > nmax ? n : nmax;
    }
    A.by = set.lv[0].border; A.bxl = set.lv[0].bxl; A.bxr = set.lv[0].border;
    dim3 grid(min((nmax / 16 + 255) / 256 + 1, 24), batch, set.top + 1);
    border_fill_kernel<<<grid, 256, 0, s>>>(A);
    PMV_LAUNCH_CHECK(ctx, "border_fill_kernel");
    return PMV_OK;
}

// d_src / d_src2: caller images for the first / second half of the batch (prev / next image sets share one
// bordered allocation so every pyramid kernel runs once per step); d_src2 == nullptr -> one source.
int pmv_internal_pyr_run(pmv_ctx *ctx, const PyrSet &set, int batch, const uint8_t *d_src, const uint8_t *d_src2,
                         int src_pitch, size_t src_stride, cudaStream_t s)
{
    const PyrLevel &l0 = set.lv[0];
    if (d_src) {
        const int half = d_src2 ? batch / 2 : batch;
        const uint8_t *srcs[2] = {d_src, d_src2};
        for (int k = 0; k < (d_src2 ? 2 : 1); k++) {
            int al = (((uintptr_t)srcs[k]) % 16 == 0) && (src_pitch % 16 == 0) && (src_stride % 16 == 0);
            dim3 grid((l0.cols + 511) / 512, (l0.rows + 7) / 8, half);
            import_kernel<<<grid, 256, 0, s>>>(srcs[k], src_pitch, src_stride, l0.rows, l0.cols,
                                               const_cast<uint8_t *>(l0.ptr) + (size_t)k * half * l0.img_stride,
                                               l0.pitch, l0.img_stride, al);
            PMV_LAUNCH_CHECK(ctx, "import_kernel");
        }
    }
    for (int l = 1; l <= set.top; l++) {
        const PyrLevel &a = set.lv[l - 1], &d = set.lv[l];
        dim3 grid((d.cols + PW_OUT - 1) / PW_OUT, (d.rows + 8 * PW_R - 1) / (8 * PW_R), batch);
        pyr_down_kernel<<<grid, 256, 0, s>>>(a.ptr, a.rows, a.cols, a.pitch, a.img_stride,
                                             const_cast<uint8_t *>(d.ptr), d.rows, d.cols, d.pitch, d.img_stride);
        PMV_LAUNCH_CHECK(ctx, "pyr_down_kernel");
    }
    return fill_borders(ctx, set, batch, s);
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
    PyrSet set;
    int rc = pmv_internal_pyr_plan(ctx, 0, 1, rows, cols, 2, win_w, win_h, max_level, &set);
    if (rc) return rc;
    const PyrLevel &l0 = set.lv[0];
    PMV_CUDA_TRY(ctx, cudaMemcpy2DAsync(const_cast<uint8_t *>(l0.ptr), l0.pitch, img, step, cols, rows,
                                        cudaMemcpyHostToDevice, ctx->stream));
    rc = pmv_internal_pyr_run(ctx, set, 1, nullptr, nullptr, 0, 0, ctx->stream);
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

}  // namespace
