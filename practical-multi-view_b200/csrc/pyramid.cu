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
//   pyr_fused_kernel   one TMA box load per 128 x 32 tile of level l -> bordered level-0 copy, Scharr plane of level l
//                      and level l + 1, all from the same shared-memory tile (one launch per level)
#include <cuda.h>

#include "common.cuh"

namespace {

// ------------------------------------------------------------------ fused level kernel ---
// One CTA = one 128 x 32 tile of source level l of one image.  The tile and its halo (2 rows above / below, 8 bytes
// left / right) arrive in shared memory through ONE TMA box load (cp.async.bulk.tensor, u8 tensor map over the
// image interior: out-of-image bytes come back as zero and are patched to BORDER_REFLECT_101 in shared memory by the
// edge tiles).  From that single pass over the source the CTA writes
//   (COPY)  the interior of the bordered level-0 copy the tracker reads          16 B stores
//   (deriv) the Scharr derivative plane of level l, one packed word per pixel    16 B stores
//   (down)  level l + 1 = pyrDown(level l), packed 16-bit SIMD-in-register       4 B stores
// so every level is read from HBM once and each output byte is written once.
constexpr int FT_W = 128, FT_H = 32, FT_HX = 16, FT_HY = 2;   // the box must START on a 16-byte boundary (tools/tma_probe.cu)
constexpr int FT_P = FT_W + 2 * FT_HX;      // 160: tile pitch = TMA box width (multiple of 16 B)
constexpr int FT_R = FT_H + 2 * FT_HY;      // 36 rows
constexpr uint32_t FT_BYTES = FT_P * FT_R;  // 5760

struct FusedArgs {
    int rows, cols;                 // source level
    int n_first;                    // images [0, n_first) come from map A, the rest from map B (caller prev / next)
    uint8_t *copy; int cpitch; size_t cstride;           // bordered level-0 interior (COPY)
    int *der; int dpitch; size_t dstride; int n_deriv;   // derivative plane of images [0, n_deriv)
    uint8_t *down; int drows, dcols, wpitch; size_t wstride;   // level l + 1 (nullptr: top level)
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <bool COPY>
__global__ void __launch_bounds__(256)
pyr_fused_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const FusedArgs A)
{
    __shared__ __align__(128) uint8_t tile[FT_BYTES];
    __shared__ __align__(8) uint64_t bar;
    const int t = threadIdx.x;
    const int x0 = blockIdx.x * FT_W, y0 = blockIdx.y * FT_H, z = blockIdx.z;
    if (t == 0) {
        const uint32_t b = smem_u32(&bar), d = smem_u32(tile);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(FT_BYTES) : "memory");
        const bool first = z < A.n_first;
        const CUtensorMap *m = first ? &mapA : &mapB;
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(d), "l"(m), "r"(x0 - FT_HX), "r"(y0 - FT_HY), "r"(first ? z : z - A.n_first), "r"(b) : "memory");
    }
    __syncthreads();   // the initialised barrier is visible to every waiter
    {
        const uint32_t b = smem_u32(&bar);
        asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@!p bra WAIT_%=;\n}" ::"r"(b) : "memory");
    }
    const int rows = A.rows, cols = A.cols;
    // edge tiles: the (at most two) rows / columns beyond each image edge that the taps reach -> reflect-101
    if (x0 == 0 || y0 == 0 || x0 + FT_W + 2 > cols || y0 + FT_H + 2 > rows) {
        for (int i = t; i < 4 * FT_P + 4 * FT_R; i += 256) {
            int r, c;   // tile coordinates
            if (i < 4 * FT_P) {
                const int k = i / FT_P;
                c = i - k * FT_P;
                const int y = k < 2 ? k - 2 : rows + (k - 2);   // image rows -2, -1, rows, rows + 1
                r = y - (y0 - FT_HY);
            } else {
                const int j = i - 4 * FT_P, k = j / FT_R;
                r = j - k * FT_R;
                const int x = k < 2 ? k - 2 : cols + (k - 2);
                c = x - (x0 - FT_HX);
            }
            if (r < 0 || r >= FT_R || c < 0 || c >= FT_P) continue;
            const int y = y0 - FT_HY + r, x = x0 - FT_HX + c;
            if (y >= 0 && y < rows && x >= 0 && x < cols) continue;
            const int sy = reflect101(y < -2 ? -2 : (y > rows + 1 ? rows + 1 : y), rows) - (y0 - FT_HY);
            const int sx = reflect101(x < -2 ? -2 : (x > cols + 1 ? cols + 1 : x), cols) - (x0 - FT_HX);
            if (sy >= 0 && sy < FT_R && sx >= 0 && sx < FT_P) tile[r * FT_P + c] = tile[sy * FT_P + sx];
        }
        __syncthreads();
    }
    if (COPY) {
        const int r = t >> 3, ch = t & 7;
        const int y = y0 + r, x = x0 + 16 * ch;
        if (y < rows && x < cols) {
            // the right border of the destination absorbs the tail of the last chunk (filled afterwards)
            *reinterpret_cast<uint4 *>(A.copy + (size_t)z * A.cstride + (size_t)y * A.cpitch + x) =
                *reinterpret_cast<const uint4 *>(tile + (r + FT_HY) * FT_P + FT_HX + 16 * ch);
        }
    }
    if (z < A.n_deriv) {
        // Scharr: 4 pixels x 4 rows per thread, three aligned words per source row (bytes x - 4 .. x + 7)
        const int strip = t & 31, rg = t >> 5;
        const int x = x0 + 4 * strip;
        if (x < cols && y0 + 4 * rg < rows) {
            const uint32_t *wp = reinterpret_cast<const uint32_t *>(tile + (4 * rg + FT_HY - 1) * FT_P + FT_HX - 4 + 4 * strip);
            uint32_t w[6][3];
#pragma unroll
            for (int r = 0; r < 6; r++)
#pragma unroll
                for (int k = 0; k < 3; k++) w[r][k] = wp[r * (FT_P / 4) + k];
            int *dp = A.der + (size_t)z * A.dstride + (size_t)(y0 + 4 * rg) * A.dpitch + x;
#pragma unroll
            for (int rr = 0; rr < 4; rr++) {
                if (y0 + 4 * rg + rr >= rows) break;
                int s0[6], s1[6];
#pragma unroll
                for (int c = 0; c < 6; c++) {
                    const int bi = c + 3;
                    const int a0 = (w[rr][bi >> 2] >> (8 * (bi & 3))) & 255, a1 = (w[rr + 1][bi >> 2] >> (8 * (bi & 3))) & 255,
                              a2 = (w[rr + 2][bi >> 2] >> (8 * (bi & 3))) & 255;
                    s0[c] = 3 * (a0 + a2) + 10 * a1;
                    s1[c] = a2 - a0;
                }
                int out[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int gx = s0[k + 2] - s0[k], gy = 3 * (s1[k] + s1[k + 2]) + 10 * s1[k + 1];
                    out[k] = (x + k < cols) ? ((gx & 0xffff) | (int)((unsigned)gy << 16)) : 0;
                }
                *reinterpret_cast<int4 *>(dp + (size_t)rr * A.dpitch) = make_int4(out[0], out[1], out[2], out[3]);
            }
        }
    }
    if (A.down) {
        // pyrDown: 4 outputs per thread; vertical [1 4 6 4 1] on packed even / odd columns, horizontal on packed pairs
        const int j = t & 15, orow = t >> 4;
        const int ox = (x0 >> 1) + 4 * j, oy = (y0 >> 1) + orow;
        if (ox < A.dcols && oy < A.drows) {
            const uint32_t M = 0x00ff00ffu;
            uint32_t VE[4], VO[4];   // the four words from source column 2 ox - 4 on (bytes 4 .. 19 of three 8-byte loads)
            {
                uint32_t e[5][4], o[5][4];
#pragma unroll
                for (int r = 0; r < 5; r++) {
                    const uint2 *sp = reinterpret_cast<const uint2 *>(tile + (2 * orow + r) * FT_P + FT_HX - 8 + 8 * j);
                    const uint2 q0 = sp[0], q1 = sp[1], q2 = sp[2];
                    const uint32_t wv[4] = {q0.y, q1.x, q1.y, q2.x};
#pragma unroll
                    for (int k = 0; k < 4; k++) { e[r][k] = wv[k] & M; o[r][k] = (wv[k] >> 8) & M; }
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    VE[k] = e[0][k] + e[4][k] + ((e[1][k] + e[3][k]) << 2) + e[2][k] * 6u;
                    VO[k] = o[0][k] + o[4][k] + ((o[1][k] + o[3][k]) << 2) + o[2][k] * 6u;
                }
            }
            // V index of word k (k = 0..3 <-> bytes 4 + 4k ..): VE[k] = (V[4+4k], V[6+4k]), VO[k] = (V[5+4k], V[7+4k]);
            // output m is centred on V[8 + 2m]
            const uint32_t A01 = __byte_perm(VE[0], VE[1], 0x5432);   // (V6,  V8)
            const uint32_t F01 = __byte_perm(VE[1], VE[2], 0x5432);   // (V10, V12)
            const uint32_t B01 = __byte_perm(VO[0], VO[1], 0x5432);   // (V7,  V9)
            const uint32_t r01 = A01 + F01 + ((B01 + VO[1]) << 2) + VE[1] * 6u + 0x00800080u;
            const uint32_t F23 = __byte_perm(VE[2], VE[3], 0x5432);   // (V14, V16)
            const uint32_t B23 = __byte_perm(VO[1], VO[2], 0x5432);   // (V11, V13)
            const uint32_t r23 = F01 + F23 + ((B23 + VO[2]) << 2) + VE[2] * 6u + 0x00800080u;
            // row pitch padding absorbs a partial word at the right edge (border fill follows)
            *reinterpret_cast<uint32_t *>(A.down + (size_t)z * A.wstride + (size_t)oy * A.wpitch + ox) = __byte_perm(r01, r23, 0x7531);
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

// ------------------------------------------------------------------ tensor maps ---------
// cuTensorMapEncodeTiled through the runtime's driver entry point (no libcuda link dependency).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

static bool tma_ok(const void *ptr, int pitch, size_t stride, int nimg)
{
    return ((uintptr_t)ptr % 16 == 0) && (pitch % 16 == 0) && (nimg == 1 || stride % 16 == 0);
}

// u8 images [nimg][rows][cols] (row pitch / image stride in bytes), box = one fused tile with its halo
static int make_image_map(pmv_ctx *ctx, CUtensorMap *m, const uint8_t *ptr, int rows, int cols, int pitch, size_t stride, int nimg)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return ctx->fail(PMV_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    if (nimg == 1 || stride % 16) stride = (size_t)pitch * (rows > 0 ? rows : 1);
    const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)nimg};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)stride};
    const cuuint32_t box[3] = {FT_P, FT_R, 1}, es[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t *>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return ctx->fail(PMV_ERR_CUDA, "cuTensorMapEncodeTiled failed");
    return PMV_OK;
}

// d_src / d_src2: caller images for the first / second half of the batch (prev / next image sets share one
// bordered allocation so every pyramid kernel runs once per step); d_src2 == nullptr -> one source;
// d_src == nullptr -> level 0 was already copied into the planned interior.  dv != nullptr: the Scharr planes of
// images [0, n_deriv) are written from the same pass.  One fused launch per level, then one border launch.
int pmv_internal_pyr_run(pmv_ctx *ctx, const PyrSet &set, int batch, const uint8_t *d_src, const uint8_t *d_src2,
                         int src_pitch, size_t src_stride, const DerivSet *dv, int n_deriv, cudaStream_t s)
{
    const PyrLevel &l0 = set.lv[0];
    const int half = d_src2 ? batch / 2 : batch;
    bool direct = d_src && tma_ok(d_src, src_pitch, src_stride, half) && (!d_src2 || tma_ok(d_src2, src_pitch, src_stride, half));
    if (d_src && !direct) {   // unaligned caller images: import first, then the fused pass reads our own level 0
        const uint8_t *srcs[2] = {d_src, d_src2};
        for (int k = 0; k < (d_src2 ? 2 : 1); k++) {
            dim3 grid((l0.cols + 511) / 512, (l0.rows + 7) / 8, half);
            import_kernel<<<grid, 256, 0, s>>>(srcs[k], src_pitch, src_stride, l0.rows, l0.cols,
                                               const_cast<uint8_t *>(l0.ptr) + (size_t)k * half * l0.img_stride,
                                               l0.pitch, l0.img_stride, 0);
            PMV_LAUNCH_CHECK(ctx, "import_kernel");
        }
    }
    if (!dv) n_deriv = 0;
    for (int l = 0; l <= set.top; l++) {
        const PyrLevel &a = set.lv[l];
        const bool down = l < set.top;
        if (!down && n_deriv == 0) break;
        FusedArgs A;
        memset(&A, 0, sizeof A);
        A.rows = a.rows; A.cols = a.cols; A.n_first = batch;
        CUtensorMap mA, mB;
        int rc;
        const bool copy = (l == 0 && direct);
        if (copy) {
            rc = make_image_map(ctx, &mA, d_src, a.rows, a.cols, src_pitch, src_stride, half);
            if (rc) return rc;
            mB = mA;
            if (d_src2) {
                rc = make_image_map(ctx, &mB, d_src2, a.rows, a.cols, src_pitch, src_stride, half);
                if (rc) return rc;
                A.n_first = half;
            }
            A.copy = const_cast<uint8_t *>(a.ptr); A.cpitch = a.pitch; A.cstride = a.img_stride;
        } else {
            rc = make_image_map(ctx, &mA, a.ptr, a.rows, a.cols, a.pitch, a.img_stride, batch);
            if (rc) return rc;
            mB = mA;
        }
        if (n_deriv > 0) {
            A.der = const_cast<int *>(dv->lv[l].ptr); A.dpitch = dv->lv[l].pitch; A.dstride = dv->lv[l].img_stride; A.n_deriv = n_deriv;
        }
        // the top level only carries derivatives: no CTA for images without them
        const int nz = down ? batch : n_deriv;
        if (down) {
            const PyrLevel &d = set.lv[l + 1];
            A.down = const_cast<uint8_t *>(d.ptr); A.drows = d.rows; A.dcols = d.cols; A.wpitch = d.pitch; A.wstride = d.img_stride;
        }
        dim3 grid((a.cols + FT_W - 1) / FT_W, (a.rows + FT_H - 1) / FT_H, nz);
        if (copy) pyr_fused_kernel<true><<<grid, 256, 0, s>>>(mA, mB, A);
        else pyr_fused_kernel<false><<<grid, 256, 0, s>>>(mA, mB, A);
        PMV_LAUNCH_CHECK(ctx, "pyr_fused_kernel");
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
    rc = pmv_internal_pyr_run(ctx, set, 1, nullptr, nullptr, 0, 0, nullptr, 0, ctx->stream);
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
