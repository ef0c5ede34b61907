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
//   pyr_fused_kernel   one TMA box load per 128 x 32 tile of level l -> bordered level-0 copy, Scharr plane of level l
//                      and level l + 1, all from the same shared-memory tile (one launch per level)
#include <cuda.h>

#include "common.cuh"

namespace {

// ------------------------------------------------------------------ fused level kernel ---
// One CTA = one 128 x 32 tile of source level l of one image.  The tile and its halo (2 rows above / below, 16 bytes
// left / right) arrive in shared memory through ONE TMA box load (cp.async.bulk.tensor, u8 tensor map over the
// image interior: out-of-image bytes come back as zero and are patched to BORDER_REFLECT_101 in shared memory by the
// edge tiles).  From that single pass over the source the CTA writes
//   (COPY)   the interior of the bordered level-0 copy the tracker reads         16 B stores
//   (deriv)  the Scharr derivative plane of level l, one packed word per pixel   16 B stores, byte dot products
//   (down)   level l + 1 = pyrDown(level l): 25 taps as ten 4-way byte dot products   4 B stores
//   (border) edge tiles: the reflect-101 border of level l next to their tile     16 B stores
// so every level is read from HBM once and each output byte is written once.
constexpr int FT_W = 128, FT_H = 32, FT_HX = 16, FT_HY = 2;   // the box must START on a 16-byte boundary (tools/tma_probe.cu)
constexpr int FT_P = FT_W + 2 * FT_HX;      // 160: tile pitch = TMA box width (multiple of 16 B)
constexpr int FT_R = FT_H + 2 * FT_HY;      // 36 rows
constexpr uint32_t FT_BYTES = FT_P * FT_R;  // 5760

struct FusedArgs {
    int rows, cols;                 // source level
    int n_first;                    // images [0, n_first) come from map A / srcA, the rest from map B / srcB (caller prev / next)
    uint8_t *copy; int cpitch; size_t cstride;           // bordered level-0 interior (COPY)
    int *der; int dpitch; size_t dstride; int n_deriv;   // derivative plane of images [0, n_deriv)
    uint8_t *down; int drows, dcols, wpitch; size_t wstride;   // level l + 1 (nullptr: top level)
    // border of level l: read from the source interior, written around `bdst` (COPY: the level-0 copy; else in place)
    const uint8_t *srcA, *srcB; int spitch; size_t sstride;
    uint8_t *bdst; int bpitch; size_t bstride;
    int by, bxl, bxr;               // border rows (top = bottom), allocated bytes on the left (multiple of 16), border columns on the right
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// reflect-101 with one fold on the fast path (the overshoot is almost always smaller than the image)
__device__ __forceinline__ int reflect101_fast(int p, int len)
{
    const int q = p < 0 ? -p : (p >= len ? 2 * len - 2 - p : p);
    return (unsigned)q < (unsigned)len ? q : reflect101(p, len);
}
__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int c)   // unsigned pixels x signed coefficients
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__host__ __device__ constexpr int pk4(int b0, int b1, int b2, int b3)
{
    return (b0 & 255) | ((b1 & 255) << 8) | ((b2 & 255) << 16) | (int)((unsigned)(b3 & 255) << 24);
}

// Can the left / right border of the rows of a tile be built from the staged tile itself?  (the mirrored columns must
// lie inside the 128 columns + 16-byte halo the tile holds)
__device__ __forceinline__ bool fused_lr_from_tile(const FusedArgs &A, int x0)
{
    const int cols = A.cols;
    const bool xlast = x0 + FT_W >= cols;
    return A.bxl + 4 <= FT_W && A.bxl + 4 <= cols && (!xlast || (cols - x0 + FT_HX >= A.bxr + 20 && A.bxr + 20 <= cols));
}

// Border of level l around one tile (edge tiles only), straight from the source in global memory (it runs while the
// tile's box load is in flight).  Part 1: rows above / below the image over the tile's image columns, whole 16-byte
// chunks.  Part 2: everything left of column 0 and from the chunk that straddles the right image edge on (that chunk
// belongs here, not to the copy phase), for all rows of the tile's outward extension.
__device__ __forceinline__ void fused_border(const FusedArgs &A, int x0, int y0, int z, bool first, int t)
{
    const int rows = A.rows, cols = A.cols;
    const bool xlast = x0 + FT_W >= cols, ylast = y0 + FT_H >= rows;
    if (x0 == 0 || y0 == 0 || xlast || ylast) {
        const uint8_t *src = (first ? A.srcA : A.srcB) + (size_t)(first ? z : z - A.n_first) * A.sstride;
        uint8_t *dst = A.bdst + (size_t)z * A.bstride;
        const int ca = cols & ~15;
        if (y0 == 0 || ylast) {
            const int nch = ((x0 + FT_W < ca ? x0 + FT_W : ca) - x0) >> 4;
            const int ntop = y0 == 0 ? A.by : 0, nrow = ntop + (ylast ? A.by : 0);
            for (int i = t; i < nrow * 8; i += 256) {
                const int k = i >> 3, ch = i & 7;
                if (ch < nch) {
                    const int y = k < ntop ? k - ntop : rows + (k - ntop);
                    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + (size_t)reflect101_fast(y, rows) * A.spitch + x0 + 16 * ch));
                    *reinterpret_cast<uint4 *>(dst + (ptrdiff_t)y * A.bpitch + x0 + 16 * ch) = v;
                }
            }
        }
        // Part 2.  Regular case (the mirrored columns lie inside this tile): the rows of the tile itself are written
        // from the staged tile by fused_tile; here only the corner rows above / below the image remain.
        const bool regular = fused_lr_from_tile(A, x0);
        int ey0 = y0 == 0 ? -A.by : y0, ey1 = ylast ? rows + A.by : y0 + FT_H;
        for (int part = 0; part < 2; part++) {
            int ya = ey0, yb = ey1;
            if (regular) {   // part 0: rows above the image, part 1: rows below
                if (part == 0) { ya = ey0; yb = ey0 < 0 ? 0 : ey0; } else { ya = ey1 > rows ? rows : ey1; yb = ey1; }
            } else if (part == 1) break;
            if (ya >= yb) continue;
            const int wl = x0 == 0 ? A.bxl >> 2 : 0;
            const int nw = wl + (xlast ? ((cols + A.bxr + 15 - ca) >> 4) << 2 : 0);
            for (int w = t & 31; w < nw; w += 32) {      // a lane keeps its word column: the four source columns are fixed
                const int xw = w < wl ? 4 * w - A.bxl : ca + 4 * (w - wl);
                const int c0 = reflect101_fast(xw, cols), c1 = reflect101_fast(xw + 1, cols), c2 = reflect101_fast(xw + 2, cols),
                          c3 = reflect101_fast(xw + 3, cols);
                for (int y = ya + (t >> 5); y < yb; y += 8) {
                    const uint8_t *srow = src + (size_t)reflect101_fast(y, rows) * A.spitch;
                    const uint32_t q = (uint32_t)__ldg(srow + c0) | ((uint32_t)__ldg(srow + c1) << 8) | ((uint32_t)__ldg(srow + c2) << 16) |
                                       ((uint32_t)__ldg(srow + c3) << 24);
                    *reinterpret_cast<uint32_t *>(dst + (ptrdiff_t)y * A.bpitch + xw) = q;
                }
            }
        }
    }
}

// Everything that is computed from one staged tile: edge patch, level-0 copy, Scharr plane, next level.
template <bool COPY>
__device__ __forceinline__ void fused_tile(const FusedArgs &A, uint8_t *tile, int x0, int y0, int z, int t)
{
    const int rows = A.rows, cols = A.cols;
    // left / right border of this tile's rows from the staged image columns (regular case; see fused_border):
    // 16-byte chunks, the mirror image of four columns is a byte permutation of two neighbouring words
    if ((x0 == 0 || x0 + FT_W >= cols) && fused_lr_from_tile(A, x0)) {
        uint8_t *dst = A.bdst + (size_t)z * A.bstride;
        const int nl = x0 == 0 ? A.bxl >> 4 : 0;                                       // chunks left of column 0
        const int ca = cols & ~15;
        const int nr = x0 + FT_W >= cols ? (cols + A.bxr + 15 - ca) >> 4 : 0;          // chunks from `ca` on
        for (int i = t; i < (nl + nr) * FT_H; i += 256) {
            const int r = i & (FT_H - 1), ch = i >> 5;
            const int y = y0 + r;
            if (y >= rows) continue;
            const uint8_t *trow = tile + (r + FT_HY) * FT_P + FT_HX - x0;              // trow[x] = image column x
            uint32_t q[4];
            int x16;
            if (ch < nl) {
                x16 = -16 * (ch + 1);                                                 // destination bytes x16 .. x16 + 15 <- columns -x16 .. -x16 - 15
                const uint32_t *w = reinterpret_cast<const uint32_t *>(trow - x16 - 16);   // w[0] = columns -x16-16 .. ; w[4] = columns -x16 ..
#pragma unroll
                for (int k = 0; k < 4; k++) q[k] = __byte_perm(w[3 - k], w[4 - k], 0x1234);
            } else {
                x16 = ca + 16 * (ch - nl);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int xw = x16 + 4 * k;
                    if (xw + 4 <= cols) {
                        q[k] = *reinterpret_cast<const uint32_t *>(trow + xw);
                    } else if (xw >= cols) {
                        const int a4 = 2 * cols - 5 - xw;                              // columns a4 .. a4 + 3, reversed
                        const uint32_t *w = reinterpret_cast<const uint32_t *>(trow + (a4 & ~3));
                        q[k] = __byte_perm(__funnelshift_r(w[0], w[1], 8 * (a4 & 3)), 0u, 0x0123);
                    } else {
                        uint32_t v = 0;
#pragma unroll
                        for (int b = 0; b < 4; b++) { const int x = xw + b; v |= (uint32_t)trow[x < cols ? x : 2 * cols - 2 - x] << (8 * b); }
                        q[k] = v;
                    }
                }
            }
            *reinterpret_cast<uint4 *>(dst + (ptrdiff_t)y * A.bpitch + x16) = make_uint4(q[0], q[1], q[2], q[3]);
        }
    }
    // edge tiles: the two rows / columns beyond each image edge that the taps reach -> reflect-101 inside the tile
    if (x0 == 0 || y0 == 0 || x0 + FT_W + 2 > cols || y0 + FT_H + 2 > rows) {
        if (t < FT_P) {   // rows -2, -1, rows, rows + 1: one tile column per thread (columns reflected too)
            const int c = t, x = x0 - FT_HX + c;
            const int sx = reflect101(x < -2 ? -2 : (x > cols + 1 ? cols + 1 : x), cols) - (x0 - FT_HX);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int y = k < 2 ? k - 2 : rows + (k - 2);
                const int r = y - (y0 - FT_HY);
                if (r >= 0 && r < FT_R) {
                    const int sy = reflect101(y, rows) - (y0 - FT_HY);
                    if (sy >= 0 && sy < FT_R && sx >= 0 && sx < FT_P) tile[r * FT_P + c] = tile[sy * FT_P + sx];
                }
            }
        }
        // columns: 4 candidate columns x 36 rows = 144 items on threads 0..143
        if (t < 4 * FT_R) {
            const int k = t / FT_R, r = t - k * FT_R;
            const int x = k < 2 ? k - 2 : cols + (k - 2);
            const int c = x - (x0 - FT_HX);
            const int y = y0 - FT_HY + r;
            if (c >= 0 && c < FT_P && y >= 0 && y < rows) {   // rows outside the image were done above (with reflected columns)
                const int sx = reflect101(x, cols) - (x0 - FT_HX);
                if (sx >= 0 && sx < FT_P) tile[r * FT_P + c] = tile[r * FT_P + sx];
            }
        }
        __syncthreads();
    }
    if (COPY) {
        const int r = t >> 3, ch = t & 7;
        const int y = y0 + r, x = x0 + 16 * ch;
        if (y < rows && x + 16 <= cols)
            *reinterpret_cast<uint4 *>(A.copy + (size_t)z * A.cstride + (size_t)y * A.cpitch + x) =
                *reinterpret_cast<const uint4 *>(tile + (r + FT_HY) * FT_P + FT_HX + 16 * ch);
    }
    if (z < A.n_deriv) {
        // Scharr: 4 pixels x 4 rows per thread.  Per source row three aligned words -> the byte windows
        // (x-1 .. x+2), (x .. x+3), (x+2 .. x+5); d/dx and d/dy are byte dot products against signed coefficient words.
        const int strip = t & 31, rg = t >> 5;
        const int x = x0 + 4 * strip;
        if (x < cols && y0 + 4 * rg < rows) {
            const uint32_t *wp = reinterpret_cast<const uint32_t *>(tile + (4 * rg + FT_HY - 1) * FT_P + FT_HX - 4 + 4 * strip);
            uint32_t Wa[6], Wc[6], Wb[6];
#pragma unroll
            for (int r = 0; r < 6; r++) {
                const uint32_t w0 = wp[r * (FT_P / 4)], w1 = wp[r * (FT_P / 4) + 1], w2 = wp[r * (FT_P / 4) + 2];
                Wa[r] = __funnelshift_r(w0, w1, 24); Wc[r] = w1; Wb[r] = __funnelshift_r(w1, w2, 16);
            }
            constexpr int X3 = pk4(-3, 0, 3, 0), X10 = pk4(-10, 0, 10, 0), YP = pk4(3, 10, 3, 0), YN = pk4(-3, -10, -3, 0);
            constexpr int X3s = pk4(0, -3, 0, 3), X10s = pk4(0, -10, 0, 10), YPs = pk4(0, 3, 10, 3), YNs = pk4(0, -3, -10, -3);
            int *dp = A.der + (size_t)z * A.dstride + (size_t)(y0 + 4 * rg) * A.dpitch + x;
#pragma unroll
            for (int rr = 0; rr < 4; rr++) {
                if (y0 + 4 * rg + rr >= rows) break;
                int gx[4], gy[4];
                gx[0] = dp4a_us(Wa[rr], X3, dp4a_us(Wa[rr + 1], X10, dp4a_us(Wa[rr + 2], X3, 0)));
                gy[0] = dp4a_us(Wa[rr + 2], YP, dp4a_us(Wa[rr], YN, 0));
                gx[1] = dp4a_us(Wc[rr], X3, dp4a_us(Wc[rr + 1], X10, dp4a_us(Wc[rr + 2], X3, 0)));
                gy[1] = dp4a_us(Wc[rr + 2], YP, dp4a_us(Wc[rr], YN, 0));
                gx[2] = dp4a_us(Wc[rr], X3s, dp4a_us(Wc[rr + 1], X10s, dp4a_us(Wc[rr + 2], X3s, 0)));
                gy[2] = dp4a_us(Wc[rr + 2], YPs, dp4a_us(Wc[rr], YNs, 0));
                gx[3] = dp4a_us(Wb[rr], X3, dp4a_us(Wb[rr + 1], X10, dp4a_us(Wb[rr + 2], X3, 0)));
                gy[3] = dp4a_us(Wb[rr + 2], YP, dp4a_us(Wb[rr], YN, 0));
                int out[4];
#pragma unroll
                for (int k = 0; k < 4; k++) out[k] = (int)__byte_perm((uint32_t)gx[k], (uint32_t)gy[k], 0x5410);
                if (x + 3 >= cols) {   // last strip of the row: the zero border of the plane starts at `cols`
#pragma unroll
                    for (int k = 0; k < 4; k++) out[k] = (x + k < cols) ? out[k] : 0;
                }
                *reinterpret_cast<int4 *>(dp + (size_t)rr * A.dpitch) = make_int4(out[0], out[1], out[2], out[3]);
            }
        }
    }
    if (A.down) {
        // pyrDown: 4 outputs per thread; every output is (sum of ten byte dot products + 128) >> 8 -- the separable
        // [1 4 6 4 1] x [1 4 6 4 1] taps with the row weight folded into the coefficient words (all exact integers)
        const int j = t & 15, orow = t >> 4;
        const int ox = (x0 >> 1) + 4 * j, oy = (y0 >> 1) + orow;
        if (ox < A.dcols && oy < A.drows) {
            uint32_t acc[4] = {128u, 128u, 128u, 128u};
#pragma unroll
            for (int r = 0; r < 5; r++) {
                const uint32_t wr = r == 2 ? 6u : ((r == 1 || r == 3) ? 4u : 1u);
                const uint32_t ca = wr * 0x04010000u, cb = wr * 0x00010406u, cc = wr * 0x04060401u, cd = wr * 0x00000001u;
                const uint2 *sp = reinterpret_cast<const uint2 *>(tile + (2 * orow + r) * FT_P + FT_HX - 8 + 8 * j);
                const uint2 q0 = sp[0], q1 = sp[1], q2 = sp[2];
                // source columns relative to 2 ox: q0.y = -4..-1, q1.x = 0..3, q1.y = 4..7, q2.x = 8..11
                acc[0] = __dp4a(q0.y, ca, __dp4a(q1.x, cb, acc[0]));   // taps -2 .. 2
                acc[1] = __dp4a(q1.x, cc, __dp4a(q1.y, cd, acc[1]));   // taps  0 .. 4
                acc[2] = __dp4a(q1.x, ca, __dp4a(q1.y, cb, acc[2]));   // taps  2 .. 6
                acc[3] = __dp4a(q1.y, cc, __dp4a(q2.x, cd, acc[3]));   // taps  4 .. 8
            }
            const uint32_t lo = __byte_perm(acc[0], acc[1], 0x0051), hi = __byte_perm(acc[2], acc[3], 0x0051);
            // row pitch padding absorbs a partial word at the right edge (the next launch writes that border)
            *reinterpret_cast<uint32_t *>(A.down + (size_t)z * A.wstride + (size_t)oy * A.wpitch + ox) = __byte_perm(lo, hi, 0x5410);
        }
    }
}

constexpr int FT_NT = 2;   // vertically adjacent tiles per CTA: both box loads are issued up front, the second lands while the first is processed

template <bool COPY>
__global__ void __launch_bounds__(256, 6)
pyr_fused_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const FusedArgs A)
{
    __shared__ __align__(128) uint8_t tile[FT_NT][FT_BYTES];
    __shared__ __align__(8) uint64_t bar[FT_NT];
    const int t = threadIdx.x;
    const int x0 = blockIdx.x * FT_W, yb = blockIdx.y * (FT_NT * FT_H), z = blockIdx.z;
    const bool first = z < A.n_first;
    if (t == 0) {
        const CUtensorMap *m = first ? &mapA : &mapB;
#pragma unroll
        for (int s = 0; s < FT_NT; s++) {
            if (yb + s * FT_H >= A.rows) break;
            const uint32_t b = smem_u32(&bar[s]), d = smem_u32(tile[s]);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(FT_BYTES) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(d), "l"(m), "r"(x0 - FT_HX), "r"(yb + s * FT_H - FT_HY), "r"(first ? z : z - A.n_first), "r"(b) : "memory");
        }
    }
    __syncthreads();   // the initialised barriers are visible to every waiter
#pragma unroll
    for (int s = 0; s < FT_NT; s++)
        if (yb + s * FT_H < A.rows) fused_border(A, x0, yb + s * FT_H, z, first, t);
#pragma unroll
    for (int s = 0; s < FT_NT; s++) {
        if (yb + s * FT_H >= A.rows) break;
        const uint32_t b = smem_u32(&bar[s]);
        asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@!p bra WAIT_%=;\n}" ::"r"(b) : "memory");
        fused_tile<COPY>(A, tile[s], x0, yb + s * FT_H, z, t);
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
    return pmv_internal_deriv_plan_buf(ctx, &ctx->deriv, &ctx->deriv_sig, set, batch, out, s);
}

int pmv_internal_deriv_plan_buf(pmv_ctx *ctx, DevBuf *buf, unsigned long long *bsig, const PyrSet &set, int batch, DerivSet *out,
                                cudaStream_t s)
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
    const void *before = buf->p;
    cudaError_t e = buf->reserve(total * 4 + 256);
    if (e != cudaSuccess) return ctx->fail(PMV_ERR_NOMEM, "derivative workspace", e);
    if (buf->p != before || *bsig != sig) {   // new geometry: (re)write the zero borders
        PMV_CUDA_TRY(ctx, cudaMemsetAsync(buf->p, 0, total * 4, s));
        *bsig = sig;
    }
    for (int l = 0; l <= set.top; l++) out->lv[l].ptr = buf->as<int>() + off[l];
    return PMV_OK;
}

// ------------------------------------------------------------------ internal planning ---
int pmv_internal_pyr_plan(pmv_ctx *ctx, int which, int batch, int rows, int cols, int border,
                          int win_w, int win_h, int max_level, PyrSet *out)
{
    return pmv_internal_pyr_plan_buf(ctx, &ctx->pyr[which], batch, rows, cols, border, win_w, win_h, max_level, out);
}

int pmv_internal_pyr_plan_buf(pmv_ctx *ctx, DevBuf *buf, int batch, int rows, int cols, int border,
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
    cudaError_t e = buf->reserve(total + 256);
    if (e != cudaSuccess) return ctx->fail(PMV_ERR_NOMEM, "pyramid workspace", e);
    for (int l = 0; l <= L; l++) out->lv[l].ptr = buf->as<uint8_t>() + off[l];
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
// images [0, n_deriv) are written from the same pass.  One fused launch per level (borders included).
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
        // border of level l: from the caller images into the level-0 copy, or in place around our own interior
        if (copy) { A.srcA = d_src; A.srcB = d_src2 ? d_src2 : d_src; A.spitch = src_pitch; A.sstride = src_stride; }
        else { A.srcA = A.srcB = a.ptr; A.spitch = a.pitch; A.sstride = a.img_stride; }
        A.bdst = const_cast<uint8_t *>(a.ptr); A.bpitch = a.pitch; A.bstride = a.img_stride;
        A.by = a.border; A.bxl = a.bxl; A.bxr = a.border;
        const int nz = batch;   // the top level runs for every image too: its border is written by this launch
        if (down) {
            const PyrLevel &d = set.lv[l + 1];
            A.down = const_cast<uint8_t *>(d.ptr); A.drows = d.rows; A.dcols = d.cols; A.wpitch = d.pitch; A.wstride = d.img_stride;
        }
        dim3 grid((a.cols + FT_W - 1) / FT_W, (a.rows + FT_NT * FT_H - 1) / (FT_NT * FT_H), nz);
        {
            ProfScope p0(l == 0 ? ctx : nullptr, PMV_PHASE_PYR_L0, s);   // the dominant launch of the group, timed on its own
            if (copy) pyr_fused_kernel<true><<<grid, 256, 0, s>>>(mA, mB, A);
            else pyr_fused_kernel<false><<<grid, 256, 0, s>>>(mA, mB, A);
        }
        PMV_LAUNCH_CHECK(ctx, "pyr_fused_kernel");
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
