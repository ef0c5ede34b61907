// lk.cu -- K3: per-feature pyramidal Lucas-Kanade solve, one warp per feature.
//
// Replaces cv::calcOpticalFlowPyrLK as called at reference OpenCVLucasKanadeFM.cpp:15
// (LKTrackerInvoker, SURVEY Appx A.3); calcSharrDeriv (Appx A.2) runs once per image level in
// scharr_level_kernel (pyramid.cu) and leaves OpenCV's derivative pyramid: packed (dI/dx | dI/dy << 16)
// per pixel, zero outside the image.
//
// Design (B200): a warp owns one feature for ALL pyramid levels (one launch per batch, no
// inter-level round trip through HBM).  Per level the warp
//   1. stages the (w+1)x(h+1) u8 patch of the previous image in shared memory (reflect-101 border
//      of the level),
//   2. stages the Scharr derivative of the same (w+1)x(h+1) bilinear support from the derivative
//      plane (computing it per feature cost 22 % of this kernel's instructions),
//   3. builds the template: 14-bit fixed-point bilinear samples of I, dI/dx, dI/dy kept in
//      REGISTERS (pixel p = lane + 32k), structure tensor by exact per-lane int32 sums +
//      REDUX warp reductions,
//   4. stages a (w+1+2M)x(h+1+2M) window of the next image and runs <= max_count Newton
//      iterations entirely out of shared memory; the mismatch vector is again an exact integer
//      sum reduced with REDUX.  The window is re-staged only if the point drifts > M px.
// All floating point that decides status / termination is fp32 with explicit round-to-nearest
// intrinsics (no FMA contraction) in OpenCV's operation order, so status flags match and
// positions agree to ~1e-4 px (gate: 0.01 px).
//
// Roofline: NOT HBM bound (1.28 MB compulsory bytes per 2 000-feature pair vs ~6e8 integer
// ops); the limiter is the integer/LSU issue rate -- see DESIGN.md.
#include <stdlib.h>

#include <vector>

#include "common.cuh"

namespace {

constexpr int LK_WARPS = 8;   // features per CTA
constexpr int LK_M = 4;       // drift margin (px) of the staged next-image window

struct LKParams {
    PyrLevel prev[PMV_MAX_PYR_LEVELS];
    PyrLevel next[PMV_MAX_PYR_LEVELS];
    DerivLevel dprev[PMV_MAX_PYR_LEVELS];   // Scharr derivative of prev (scharr_level_kernel), zero outside the image
    int top;
    const float *prev_xy;
    float *next_xy;
    uint8_t *status;
    float *err;
    int n;
    int win_w, win_h;
    int max_count;
    double eps2;
    int flags;
    float min_eig;
    int smem_per_warp;
    int pp;  // byte pitch of the staged prev patch (multiple of 4)
    int jp;  // byte pitch of the staged next window (multiple of 4)
};

__device__ __forceinline__ long long warp_sum_i64(int v)
{
    // exact sum of 32 int32 values: REDUX on the two 16-bit halves
    int hi = __reduce_add_sync(0xffffffffu, v >> 16);
    int lo = __reduce_add_sync(0xffffffffu, v & 0xffff);
    return ((long long)hi << 16) + (long long)lo;
}

__device__ __forceinline__ void bilinear_weights(float a, float b, int &iw00, int &iw01, int &iw10, int &iw11)
{
    const float s = 16384.f;
    float na = __fsub_rn(1.f, a), nb = __fsub_rn(1.f, b);
    iw00 = __float2int_rn(__fmul_rn(__fmul_rn(na, nb), s));
    iw01 = __float2int_rn(__fmul_rn(__fmul_rn(a, nb), s));
    iw10 = __float2int_rn(__fmul_rn(__fmul_rn(na, b), s));
    iw11 = 16384 - iw00 - iw01 - iw10;
}

// ((sum_i q_i * w_i + 256) >> 9) - ival with the four u8 neighbours packed in `quad`, the weights split
// into signed high / unsigned low bytes and Iq = 512 * ival - 256.
__device__ __forceinline__ int lk_interp_diff(unsigned quad, unsigned wlo, unsigned whi, int Iq)
{
    int t;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(t) : "r"(quad), "r"(whi), "r"(0));
    const unsigned u = __dp4a(quad, wlo, (unsigned)(-Iq));
    return (int)(((unsigned)t << 8) + u) >> 9;
}

// Copy the rows [y0, y0+nrows) x bytes [x0, x0+nbytes) of a bordered level into shared memory with
// aligned 4-byte asynchronous copies.  The smem copy keeps the global misalignment: pixel (r, c) of the region is
// at dst[r*spitch + mis + c] with mis = x0 & 3 (returned).  Level interiors are 16 B aligned and
// pitched to 128 B, so the misalignment is the same for every row.
__device__ __forceinline__ void cp_async_4(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// The copies are ASYNCHRONOUS (cp.async, LDGSTS): global -> shared memory without a register round trip, so the
// issuing warp does not stall per word; the caller waits once (cp_async_wait_all + __syncwarp) before reading.
__device__ __forceinline__ int stage_region(uint32_t *dst, int spitch_words, const uint8_t *img, int pitch,
                                            int x0, int y0, int nbytes, int nrows, int lane)
{
    const int mis = x0 & 3;
    const int nw = (mis + nbytes + 3) >> 2;
    const uint32_t *g = reinterpret_cast<const uint32_t *>(img + (ptrdiff_t)y0 * pitch + (x0 - mis));
    const int gp = pitch >> 2;
    const int total = nw * nrows;
    const int qd = 32 / nw, rm = 32 - qd * nw;
    int r = lane / nw, cw = lane - r * nw;
    for (int i = lane; i < total; i += 32) {
        cp_async_4(dst + r * spitch_words + cw, g + r * gp + cw);
        cw += rm; r += qd;
        if (cw >= nw) { cw -= nw; r++; }
    }
    return mis;
}

// The dispatch picks the smallest KPIX in {4, 8, 11, 14, 17, 20, 24, 28, 32} that covers the window, so the
// window has more than 32 * lk_kprev(KPIX) pixels: template slots k < lk_kprev(KPIX) hold a pixel in every lane.
__host__ __device__ constexpr int lk_kprev(int kpix)
{
    return kpix <= 4 ? 0 : kpix <= 8 ? 4 : kpix <= 11 ? 8 : kpix <= 14 ? 11 : kpix <= 17 ? 14 : kpix <= 20 ? 17 : kpix <= 24 ? 20 : kpix <= 28 ? 24 : 28;
}

template <int KPIX>
__global__ void __launch_bounds__(LK_WARPS * 32)
lk_track_kernel(const LKParams P)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ uint16_t s_tab[32 * KPIX];  // pixel p -> (y << 8) | x ; padding -> (0,0), masked by `valid`

    const int w = P.win_w, h = P.win_h, npx = w * h;
    for (int p = threadIdx.x; p < 32 * KPIX; p += LK_WARPS * 32) {
        int y = p / w, x = p - y * w;
        s_tab[p] = (uint16_t)(p < npx ? ((y << 8) | x) : 0);
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.x * LK_WARPS + warp;
    const int b = blockIdx.y;
    if (f >= P.n) return;  // warp-uniform; only __syncwarp below

    uint8_t *ws = smem + (size_t)warp * P.smem_per_warp;
    // one pitch for all three staged buffers (jp bytes for the patches, jp words for the derivatives): pixel
    // (y, x) sits at offset joff[k] = y * jp + x in each of them, so the template phase needs no table lookups
    const int pp = P.jp, jp = P.jp, dp = P.jp;
    uint8_t *ps = ws;                                            // prev patch (h+1) rows x jp bytes
    int *ds = reinterpret_cast<int *>(ws + (h + 1) * pp);        // derivs (h+1) rows x jp words, short2 packed
    uint8_t *js = ws;                                            // next window, aliases ps/ds
    const int jw = w + 1 + 2 * LK_M, jh = h + 1 + 2 * LK_M;

    const size_t pt = (size_t)b * P.n + f;
    const float pt_x = P.prev_xy[2 * pt], pt_y = P.prev_xy[2 * pt + 1];
    float cur_x = 0.f, cur_y = 0.f;  // nextPts[i]
    if (P.flags & PMV_LK_USE_INITIAL_FLOW) { cur_x = P.next_xy[2 * pt]; cur_y = P.next_xy[2 * pt + 1]; }
    int status = 1;
    float errv = 0.f;

    const float halfx = (w - 1) * 0.5f, halfy = (h - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);

    // per-lane template (pixel p = lane + 32k): pre-biased intensity, packed (dI/dx | dI/dy << 16), window
    // offset, and the cached 2x2 neighbourhood of the next image (J00 | J01<<8 | J10<<16 | J11<<24)
    int Iq[KPIX], dxy[KPIX], joff[KPIX];
    unsigned quad[KPIX];
    unsigned valid = 0;
#pragma unroll
    for (int k = 0; k < KPIX; k++) {
        const int t = s_tab[lane + 32 * k];
        if (lane + 32 * k < npx) valid |= 1u << k;
        joff[k] = (t >> 8) * jp + (t & 255);
    }

    for (int level = P.top; level >= 0; level--) {
        const PyrLevel &I = P.prev[level];
        const PyrLevel &J = P.next[level];
        const uint8_t *Iimg = I.ptr + (size_t)b * I.img_stride;
        const uint8_t *Jimg = J.ptr + (size_t)b * J.img_stride;
        const float sc = 1.f / (float)(1 << level);
        float px = __fmul_rn(pt_x, sc), py = __fmul_rn(pt_y, sc);
        float nx, ny;
        if (level == P.top) {
            if (P.flags & PMV_LK_USE_INITIAL_FLOW) { nx = __fmul_rn(cur_x, sc); ny = __fmul_rn(cur_y, sc); }
            else { nx = px; ny = py; }
        } else {
            nx = __fmul_rn(cur_x, 2.f); ny = __fmul_rn(cur_y, 2.f);
        }
        cur_x = nx; cur_y = ny;

        px = __fsub_rn(px, halfx); py = __fsub_rn(py, halfy);
        const int ipx = __float2int_rd(px), ipy = __float2int_rd(py);
        if (ipx < -w || ipx >= I.cols || ipy < -h || ipy >= I.rows) {
            if (level == 0) { status = 0; errv = 0.f; }
            continue;
        }
        int iw00, iw01, iw10, iw11;
        bilinear_weights(__fsub_rn(px, (float)ipx), __fsub_rn(py, (float)ipy), iw00, iw01, iw10, iw11);

        // ---- 1. stage the bilinear support of the template: prev patch and its Scharr derivative, rows
        //         ipy.., cols ipx.. ((h+1) x (w+1)).  The derivative comes from the per-level plane built once
        //         per image (scharr_level_kernel) -- as OpenCV's derivative pyramid, zero outside the image;
        //         computing it here per feature cost 22 % of this kernel's instructions.
        __syncwarp();
        const int pmis = stage_region(reinterpret_cast<uint32_t *>(ps), pp >> 2, Iimg, I.pitch, ipx, ipy, w + 1, h + 1, lane);
        {
            const DerivLevel &DL = P.dprev[level];
            const int *dsrc = DL.ptr + (size_t)b * DL.img_stride + (ptrdiff_t)ipy * DL.pitch + ipx;
            for (int c0 = 0; c0 <= w; c0 += 32) {
                const int pc = c0 + lane;
                if (pc <= w) {
                    // one asynchronous 4-byte copy per word (LDGSTS): no register round trip, one wait after the loop
                    const int *src = dsrc + pc;
                    int *dst = ds + pc;
                    for (int r = 0; r <= h; r++) { cp_async_4(dst, src); src += DL.pitch; dst += dp; }
                }
            }
        }
        cp_async_wait_all();
        __syncwarp();
        // ---- 3. template + structure tensor --------------------------------------------
        int sA11 = 0, sA12 = 0, sA22 = 0;
#pragma unroll
        for (int k = 0; k < KPIX; k++) {
            const uint8_t *q = ps + pmis + joff[k];
            int ival = (q[0] * iw00 + q[1] * iw01 + q[pp] * iw10 + q[pp + 1] * iw11 + (1 << 8)) >> 9;
            const int *d = ds + joff[k];
            int d00 = d[0], d01 = d[1], d10 = d[dp], d11 = d[dp + 1];
            int ix = ((int)(short)d00 * iw00 + (int)(short)d01 * iw01 + (int)(short)d10 * iw10 +
                      (int)(short)d11 * iw11 + (1 << 13)) >> 14;
            int iy = ((d00 >> 16) * iw00 + (d01 >> 16) * iw01 + (d10 >> 16) * iw10 +
                      (d11 >> 16) * iw11 + (1 << 13)) >> 14;
            if (k >= lk_kprev(KPIX) && !((valid >> k) & 1)) { ival = 0; ix = 0; iy = 0; }   // padding exists only past the previous KPIX step
            Iq[k] = 512 * ival - 256;                               // ((v + 256) >> 9) - ival == (v - Iq) >> 9
            dxy[k] = (ix & 0xffff) | (int)((unsigned)iy << 16);
            sA11 += ix * ix; sA12 += ix * iy; sA22 += iy * iy;
        }
        const float A11 = __fmul_rn((float)warp_sum_i64(sA11), FLT_SCALE);
        const float A12 = __fmul_rn((float)warp_sum_i64(sA12), FLT_SCALE);
        const float A22 = __fmul_rn((float)warp_sum_i64(sA22), FLT_SCALE);
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dA = __fsub_rn(A11, A22);
        const float minEig = __fdiv_rn(
            __fsub_rn(__fadd_rn(A22, A11),
                      __fsqrt_rn(__fadd_rn(__fmul_rn(dA, dA), __fmul_rn(__fmul_rn(4.f, A12), A12)))),
            (float)(2 * w * h));
        if (P.flags & PMV_LK_GET_MIN_EIGENVALS) errv = minEig;
        if (minEig < P.min_eig || D < 1.1920928955078125e-07f) {
            if (level == 0) status = 0;
            continue;
        }
        D = __fdiv_rn(1.f, D);
        nx = __fsub_rn(nx, halfx); ny = __fsub_rn(ny, halfy);

        // ---- 4. Newton iterations out of the staged next-image window --------------------
        __syncwarp();  // template reads of ps/ds are done before js overwrites them
        int jx0 = 0, jy0 = 0, jmis = 0;
        bool staged = false;
        int qx = 0, qy = 0;           // integer position the cached quads belong to
        bool have_quads = false;
        float pdx = 0.f, pdy = 0.f;
        // (Re)load the 2x2 neighbourhoods for integer position (ix_, iy_); restages the window if needed.
        auto load_quads = [&](int ix_, int iy_) {
            int ox = ix_ - jx0, oy = iy_ - jy0;
            if (!staged || ox < 0 || ox > 2 * LK_M || oy < 0 || oy > 2 * LK_M) {
                __syncwarp();
                jx0 = ix_ - LK_M; jy0 = iy_ - LK_M;
                jmis = stage_region(reinterpret_cast<uint32_t *>(js), jp >> 2, Jimg, J.pitch, jx0, jy0, jw, jh, lane);
                cp_async_wait_all();
                __syncwarp();
                staged = true;
                ox = LK_M; oy = LK_M;
            }
            const uint8_t *jb = js + oy * jp + ox + jmis;
#pragma unroll
            for (int k = 0; k < KPIX; k++) {
                const uint8_t *q = jb + joff[k];
                const unsigned lo = __byte_perm((unsigned)q[0], (unsigned)q[1], 0x0040);
                const unsigned hi = __byte_perm((unsigned)q[jp], (unsigned)q[jp + 1], 0x0040);
                quad[k] = __byte_perm(lo, hi, 0x5410);
            }
            qx = ix_; qy = iy_; have_quads = true;
        };
        // the fixed-point weights (iw11 can be -1 after rounding, iw00 can be 16384) are split as
        // w = 256 * hi + lo with hi in [-1, 64] (signed bytes) and lo in [0, 255] (unsigned bytes), so the
        // interpolation is two 4-way byte dot products: dp4a.u32.s32 (hi) and dp4a.u32.u32 (lo)
        auto split_weights = [&](unsigned &wlo, unsigned &whi) {
            wlo = (unsigned)(iw00 & 255) | ((unsigned)(iw01 & 255) << 8) | ((unsigned)(iw10 & 255) << 16) | ((unsigned)(iw11 & 255) << 24);
            whi = (unsigned)((iw00 >> 8) & 255) | ((unsigned)((iw01 >> 8) & 255) << 8) | ((unsigned)((iw10 >> 8) & 255) << 16) |
                  ((unsigned)((iw11 >> 8) & 255) << 24);
        };
        for (int j = 0; j < P.max_count; j++) {
            const int inx = __float2int_rd(nx), iny = __float2int_rd(ny);
            if (inx < -w || inx >= J.cols || iny < -h || iny >= J.rows) {
                if (level == 0) status = 0;
                break;
            }
            if (!have_quads || inx != qx || iny != qy) load_quads(inx, iny);
            bilinear_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), iw00, iw01, iw10, iw11);
            unsigned wlo, whi;
            split_weights(wlo, whi);
            int ib1 = 0, ib2 = 0;
#pragma unroll
            for (int k = 0; k < KPIX; k++) {
                const int diff = lk_interp_diff(quad[k], wlo, whi, Iq[k]);
                ib1 += diff * (int)(short)dxy[k];
                ib2 += diff * (dxy[k] >> 16);
            }
            const float b1 = __fmul_rn((float)warp_sum_i64(ib1), FLT_SCALE);
            const float b2 = __fmul_rn((float)warp_sum_i64(ib2), FLT_SCALE);
            const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
            const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
            nx = __fadd_rn(nx, dx); ny = __fadd_rn(ny, dy);
            cur_x = __fadd_rn(nx, halfx); cur_y = __fadd_rn(ny, halfy);
            if ((double)dx * (double)dx + (double)dy * (double)dy <= P.eps2) break;
            if (j > 0 && fabsf(__fadd_rn(dx, pdx)) < 0.01f && fabsf(__fadd_rn(dy, pdy)) < 0.01f) {
                cur_x = __fsub_rn(cur_x, __fmul_rn(dx, 0.5f));
                cur_y = __fsub_rn(cur_y, __fmul_rn(dy, 0.5f));
                break;
            }
            pdx = dx; pdy = dy;
        }

        // ---- 5. level-0 residual error ----------------------------------------------------
        if (status && level == 0 && !(P.flags & PMV_LK_GET_MIN_EIGENVALS)) {
            const float fx = __fsub_rn(cur_x, halfx), fy = __fsub_rn(cur_y, halfy);
            const int inx = __float2int_rd(fx), iny = __float2int_rd(fy);
            if (inx < -w || inx >= J.cols || iny < -h || iny >= J.rows) {
                status = 0;
                continue;
            }
            if (!have_quads || inx != qx || iny != qy) load_quads(inx, iny);
            bilinear_weights(__fsub_rn(fx, (float)inx), __fsub_rn(fy, (float)iny), iw00, iw01, iw10, iw11);
            unsigned wlo, whi;
            split_weights(wlo, whi);
            int e = 0;
#pragma unroll
            for (int k = 0; k < KPIX; k++) {
                const int diff = lk_interp_diff(quad[k], wlo, whi, Iq[k]);
                if ((valid >> k) & 1) e += abs(diff);
            }
            e = __reduce_add_sync(0xffffffffu, e);
            errv = __fdiv_rn((float)e, (float)(32 * w * h));
        }
    }

    if (lane == 0) {
        P.next_xy[2 * pt] = cur_x;
        P.next_xy[2 * pt + 1] = cur_y;
        P.status[pt] = (uint8_t)status;
        P.err[pt] = errv;
    }
}

template <int KPIX>
int launch_lk(pmv_ctx *ctx, LKParams &P, int batch, cudaStream_t s)
{
    size_t smem = (size_t)P.smem_per_warp * LK_WARPS;
    if (ctx->attr_first(PMV_ATTR_LK_BASE + KPIX))
        cudaFuncSetAttribute(lk_track_kernel<KPIX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    dim3 grid((P.n + LK_WARPS - 1) / LK_WARPS, batch);
    lk_track_kernel<KPIX><<<grid, LK_WARPS * 32, smem, s>>>(P);
    PMV_LAUNCH_CHECK(ctx, "lk_track_kernel");
    return PMV_OK;
}

int check_lk_args(pmv_ctx *ctx, int rows, int cols, int step, int n, int win_w, int win_h, int max_level)
{
    if (rows <= 0 || cols <= 0 || step < cols || n < 0 || max_level < 0)
        return ctx->fail(PMV_ERR_INVALID, "lk: bad argument");
    if (win_w <= 2 || win_h <= 2)  // OpenCV asserts winSize > 2
        return ctx->fail(PMV_ERR_INVALID, "lk: window must be larger than 2x2");
    if (win_w > 255 || win_h > 255 || win_w * win_h > 1024)
        return ctx->fail(PMV_ERR_UNSUPPORTED, "lk: window area above 1024 pixels is not supported");
    if (max_level >= PMV_MAX_PYR_LEVELS)
        return ctx->fail(PMV_ERR_UNSUPPORTED, "lk: max_level above 7 is not supported");
    return PMV_OK;
}

// Enqueue import + pyramids + tracking for `batch` pairs.  d_prev/d_next: device images (any pitch) or
// nullptr when the caller already copied level 0 into the planned interior (host upload path).
// prev and next image sets share ONE bordered allocation of 2*batch images (prev = [0, batch),
// next = [batch, 2*batch)), so import / pyrDown / border fill run once per step for both.
int lk_plan(pmv_ctx *ctx, int batch, int rows, int cols, int win_w, int win_h, int max_level,
            PyrSet *sp, PyrSet *sn, DerivSet *dv, cudaStream_t s)
{
    const int border = pmv_internal_lk_border(win_w, win_h);
    int rc = pmv_internal_pyr_plan(ctx, 0, 2 * batch, rows, cols, border, win_w, win_h, max_level, sp);
    if (rc) return rc;
    *sn = *sp;
    for (int l = 0; l <= sp->top; l++) sn->lv[l].ptr = sp->lv[l].ptr + (size_t)batch * sp->lv[l].img_stride;
    return pmv_internal_deriv_plan(ctx, *sp, batch, dv, s);
}

int lk_enqueue(pmv_ctx *ctx, const PyrSet &sp, const PyrSet &sn, const DerivSet &dv, const uint8_t *d_prev, const uint8_t *d_next,
               int batch, size_t img_stride, int pitch, const float *d_prev_xy, int n, int win_w, int win_h,
               int max_count, double eps, int flags, double min_eig_thr,
               float *d_next_xy, uint8_t *d_status, float *d_err, cudaStream_t s)
{
    int rc;
    {
        ProfScope ps(ctx, PMV_PHASE_PYRAMID, s);
        const DerivSet *dp = n > 0 ? &dv : nullptr;
        if (sn.lv[0].ptr == sp.lv[0].ptr + (size_t)batch * sp.lv[0].img_stride) {
            rc = pmv_internal_pyr_run(ctx, sp, 2 * batch, d_prev, d_next, pitch, img_stride, dp, batch, s);
            if (rc) return rc;
        } else {  // chunk smaller than the planned batch: the two sets are not adjacent -> two passes
            rc = pmv_internal_pyr_run(ctx, sp, batch, d_prev, nullptr, pitch, img_stride, dp, batch, s);
            if (rc) return rc;
            rc = pmv_internal_pyr_run(ctx, sn, batch, d_next, nullptr, pitch, img_stride, nullptr, 0, s);
            if (rc) return rc;
        }
    }
    return pmv_internal_lk_launch(ctx, sp, sn, dv, batch, d_prev_xy, n, win_w, win_h, max_count, eps, flags, min_eig_thr,
                                  d_next_xy, d_status, d_err, s);
}

}  // namespace

int pmv_internal_lk_border(int win_w, int win_h) { return (win_w > win_h ? win_w : win_h) + LK_M + 2; }

int pmv_internal_lk_launch(pmv_ctx *ctx, const PyrSet &sp, const PyrSet &sn, const DerivSet &dv, int batch,
                           const float *d_prev_xy, int n, int win_w, int win_h, int max_count, double eps, int flags,
                           double min_eig_thr, float *d_next_xy, uint8_t *d_status, float *d_err, cudaStream_t s)
{
    if (n == 0) return PMV_OK;
    ProfScope pl(ctx, PMV_PHASE_LK, s);
    LKParams P;
    memset(&P, 0, sizeof P);
    for (int l = 0; l <= sp.top; l++) { P.prev[l] = sp.lv[l]; P.next[l] = sn.lv[l]; P.dprev[l] = dv.lv[l]; }
    P.top = sp.top;
    P.prev_xy = d_prev_xy; P.next_xy = d_next_xy; P.status = d_status; P.err = d_err;
    P.n = n; P.win_w = win_w; P.win_h = win_h;
    if (max_count < 0) max_count = 0;
    if (max_count > 100) max_count = 100;
    if (eps < 0) eps = 0;
    if (eps > 10) eps = 10;
    P.max_count = max_count; P.eps2 = eps * eps; P.flags = flags; P.min_eig = (float)min_eig_thr;
    P.pp = align_up(3 + win_w + 1 + 3, 4);
    P.jp = align_up(3 + win_w + 1 + 2 * LK_M + 3, 4);
    int tmpl = (win_h + 1) * P.jp * 5;   // prev patch (bytes) + derivative support (words), both with pitch jp
    int jwin = (win_h + 1 + 2 * LK_M) * P.jp;
    P.smem_per_warp = align_up(tmpl > jwin ? tmpl : jwin, 16);
    const int kmax = (win_w * win_h + 31) / 32;
    if (kmax <= 4) return launch_lk<4>(ctx, P, batch, s);
    if (kmax <= 8) return launch_lk<8>(ctx, P, batch, s);
    if (kmax <= 11) return launch_lk<11>(ctx, P, batch, s);
    if (kmax <= 14) return launch_lk<14>(ctx, P, batch, s);
    if (kmax <= 17) return launch_lk<17>(ctx, P, batch, s);
    if (kmax <= 20) return launch_lk<20>(ctx, P, batch, s);
    if (kmax <= 24) return launch_lk<24>(ctx, P, batch, s);
    if (kmax <= 28) return launch_lk<28>(ctx, P, batch, s);
    return launch_lk<32>(ctx, P, batch, s);
}

extern "C" {

PMV_API int pmv_lk_track_batched_dev(pmv_ctx *ctx, const uint8_t *d_prev, const uint8_t *d_next,
                                     int batch, size_t img_stride, int rows, int cols, int step,
                                     const float *d_prev_xy, int n, int win_w, int win_h,
                                     int max_level, int max_count, double eps, int flags,
                                     double min_eig_thr, float *d_next_xy, uint8_t *d_status, float *d_err)
{
    if (!ctx) return PMV_ERR_INVALID;
    if (!d_prev || !d_next || batch <= 0 || (n > 0 && (!d_prev_xy || !d_next_xy || !d_status || !d_err)))
        return ctx->fail(PMV_ERR_INVALID, "pmv_lk_track_batched_dev: null pointer / empty batch");
    int rc = check_lk_args(ctx, rows, cols, step, n, win_w, win_h, max_level);
    if (rc) return rc;
    if (img_stride < (size_t)rows * step) return ctx->fail(PMV_ERR_INVALID, "lk: img_stride < rows*step");
    cudaSetDevice(ctx->device);
    PyrSet sp, sn;
    DerivSet dv;
    rc = lk_plan(ctx, batch, rows, cols, win_w, win_h, max_level, &sp, &sn, &dv, ctx->stream);
    if (rc) return rc;
    return lk_enqueue(ctx, sp, sn, dv, d_prev, d_next, batch, img_stride, step, d_prev_xy, n, win_w, win_h,
                      max_count, eps, flags, min_eig_thr, d_next_xy, d_status, d_err, ctx->stream);
}

PMV_API int pmv_lk_track_batched(pmv_ctx *ctx, const uint8_t *prev, const uint8_t *next, int batch,
                                 size_t img_stride, int rows, int cols, int step,
                                 const float *prev_xy, int n, int win_w, int win_h, int max_level,
                                 int max_count, double eps, int flags, double min_eig_thr,
                                 float *next_xy, uint8_t *status, float *err)
{
    if (!ctx) return PMV_ERR_INVALID;
    if (!prev || !next || batch <= 0 || (n > 0 && (!prev_xy || !next_xy || !status || !err)))
        return ctx->fail(PMV_ERR_INVALID, "pmv_lk_track_batched: null pointer / empty batch");
    int rc = check_lk_args(ctx, rows, cols, step, n, win_w, win_h, max_level);
    if (rc) return rc;
    if (img_stride < (size_t)rows * step) return ctx->fail(PMV_ERR_INVALID, "lk: img_stride < rows*step");
    cudaSetDevice(ctx->device);

    // Chunked pipeline: raw image bytes go up on the copy stream (one contiguous transfer per chunk and image set);
    // chunks alternate between TWO compute streams, each with its own pyramid / derivative slot, so that the pyramid
    // build of chunk c+1 and the tail of chunk c's tracking kernel (a few long-running features on a mostly idle GPU)
    // overlap; results go back per chunk on a fourth stream (uploads and downloads use different copy engines).
    // The first chunk is small (nothing overlaps its upload), the following ones grow to CH.
    const char *ch_env = getenv("PMV_LK_CHUNK");   // tuning knob (tools/): images per upload chunk
    const int CH = batch <= 8 ? batch : (ch_env && atoi(ch_env) > 0 ? atoi(ch_env) : 32);
    std::vector<int> chunk_begin;
    for (int b0 = 0, sz = CH < 8 ? CH : 8; b0 < batch;) {
        chunk_begin.push_back(b0);
        b0 += sz < batch - b0 ? sz : batch - b0;
        sz = 2 * sz < CH ? 2 * sz : CH;
    }
    chunk_begin.push_back(batch);
    const int nchunks = (int)chunk_begin.size() - 1;
    const int nslots = nchunks > 1 && !getenv("PMV_LK_ONE_STREAM") ? 2 : 1;
    cudaStream_t s = ctx->stream;
    // slot k: prev images [2k CH, (2k+1) CH), next images [(2k+1) CH, (2k+2) CH) of one pyramid plan; derivatives likewise
    PyrSet all, sp[2], sn[2];
    DerivSet dall, dv[2];
    rc = pmv_internal_pyr_plan(ctx, 0, 2 * nslots * CH, rows, cols, pmv_internal_lk_border(win_w, win_h), win_w, win_h, max_level, &all);
    if (rc) return rc;
    rc = pmv_internal_deriv_plan(ctx, all, nslots * CH, &dall, s);
    if (rc) return rc;
    for (int k = 0; k < nslots; k++) {
        sp[k] = all; sn[k] = all; dv[k] = dall;
        for (int l = 0; l <= all.top; l++) {
            sp[k].lv[l].ptr = all.lv[l].ptr + (size_t)(2 * k) * CH * all.lv[l].img_stride;
            sn[k].lv[l].ptr = all.lv[l].ptr + (size_t)(2 * k + 1) * CH * all.lv[l].img_stride;
            dv[k].lv[l].ptr = dall.lv[l].ptr + (size_t)k * CH * dall.lv[l].img_stride;
        }
    }
    const size_t raw_bytes = (size_t)batch * img_stride;
    cudaError_t e = ctx->pts[0].reserve((size_t)batch * n * 8 + 8);
    if (e == cudaSuccess) e = ctx->pts[1].reserve((size_t)batch * n * 8 + 8);
    if (e == cudaSuccess) e = ctx->pts[2].reserve((size_t)batch * n + 8);
    if (e == cudaSuccess) e = ctx->pts[3].reserve((size_t)batch * n * 4 + 8);
    if (e == cudaSuccess) e = ctx->img[0].reserve(raw_bytes + 16);
    if (e == cudaSuccess) e = ctx->img[1].reserve(raw_bytes + 16);
    if (e != cudaSuccess) return ctx->fail(PMV_ERR_NOMEM, "lk batch buffers", e);
    float *dpx = ctx->pts[0].as<float>(), *dnx = ctx->pts[1].as<float>();
    uint8_t *dst = ctx->pts[2].as<uint8_t>();
    float *der = ctx->pts[3].as<float>();
    uint8_t *rawP = ctx->img[0].as<uint8_t>(), *rawN = ctx->img[1].as<uint8_t>();
    cudaStream_t cs = nchunks > 1 ? ctx->copy_stream : s;
    if (nchunks > 1 && !ctx->d2h_stream) PMV_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
    if (nslots > 1 && !ctx->aux_stream) PMV_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
    cudaStream_t ds = nchunks > 1 ? ctx->d2h_stream : s;
    cudaStream_t comp[2] = {s, nslots > 1 ? ctx->aux_stream : s};
    while ((int)ctx->chunk_ev.size() < 2 * nchunks + 2) {
        cudaEvent_t ev = nullptr;
        PMV_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        ctx->chunk_ev.push_back(ev);
    }
    cudaEvent_t ev_start = ctx->chunk_ev[2 * nchunks], ev_join = ctx->chunk_ev[2 * nchunks + 1];
    if (n > 0) {
        PMV_CUDA_TRY(ctx, cudaMemcpyAsync(dpx, prev_xy, (size_t)batch * n * 8, cudaMemcpyHostToDevice, s));
        if (flags & PMV_LK_USE_INITIAL_FLOW)
            PMV_CUDA_TRY(ctx, cudaMemcpyAsync(dnx, next_xy, (size_t)batch * n * 8, cudaMemcpyHostToDevice, s));
    }
    if (nchunks > 1) {   // the side streams start behind what the caller's stream holds (points, zeroed derivative borders)
        PMV_CUDA_TRY(ctx, cudaEventRecord(ev_start, s));
        PMV_CUDA_TRY(ctx, cudaStreamWaitEvent(cs, ev_start, 0));
        if (nslots > 1) PMV_CUDA_TRY(ctx, cudaStreamWaitEvent(comp[1], ev_start, 0));
    }
    for (int c = 0; c < nchunks; c++) {
        const int b0 = chunk_begin[c], nb = chunk_begin[c + 1] - b0, k = c % nslots;
        const size_t off = (size_t)b0 * img_stride;
        // the last image of the batch may be shorter than img_stride in the caller's buffer
        const size_t bytes = (size_t)(nb - 1) * img_stride + (size_t)(rows - 1) * step + cols;
        PMV_CUDA_TRY(ctx, cudaMemcpyAsync(rawP + off, prev + off, bytes, cudaMemcpyHostToDevice, cs));
        PMV_CUDA_TRY(ctx, cudaMemcpyAsync(rawN + off, next + off, bytes, cudaMemcpyHostToDevice, cs));
        if (nchunks > 1) {
            PMV_CUDA_TRY(ctx, cudaEventRecord(ctx->chunk_ev[c], cs));
            PMV_CUDA_TRY(ctx, cudaStreamWaitEvent(comp[k], ctx->chunk_ev[c], 0));
        }
        rc = lk_enqueue(ctx, sp[k], sn[k], dv[k], rawP + off, rawN + off, nb, img_stride, step, dpx + (size_t)b0 * n * 2, n, win_w,
                        win_h, max_count, eps, flags, min_eig_thr, dnx + (size_t)b0 * n * 2, dst + (size_t)b0 * n,
                        der + (size_t)b0 * n, comp[k]);
        if (rc) return rc;
        if (n > 0) {
            if (nchunks > 1) {
                PMV_CUDA_TRY(ctx, cudaEventRecord(ctx->chunk_ev[nchunks + c], comp[k]));
                PMV_CUDA_TRY(ctx, cudaStreamWaitEvent(ds, ctx->chunk_ev[nchunks + c], 0));
            }
            const size_t o = (size_t)b0 * n, cnt = (size_t)nb * n;
            PMV_CUDA_TRY(ctx, cudaMemcpyAsync(next_xy + 2 * o, dnx + 2 * o, cnt * 8, cudaMemcpyDeviceToHost, ds));
            PMV_CUDA_TRY(ctx, cudaMemcpyAsync(status + o, dst + o, cnt, cudaMemcpyDeviceToHost, ds));
            PMV_CUDA_TRY(ctx, cudaMemcpyAsync(err + o, der + o, cnt * 4, cudaMemcpyDeviceToHost, ds));
        }
    }
    if (nchunks > 1) {   // join: the caller's stream is complete when the last download and both compute streams are
        PMV_CUDA_TRY(ctx, cudaEventRecord(ev_join, ds));
        PMV_CUDA_TRY(ctx, cudaStreamWaitEvent(s, ev_join, 0));
        if (nslots > 1) {
            PMV_CUDA_TRY(ctx, cudaEventRecord(ev_start, comp[1]));
            PMV_CUDA_TRY(ctx, cudaStreamWaitEvent(s, ev_start, 0));
        }
    }
    PMV_CUDA_TRY(ctx, cudaStreamSynchronize(s));
    return PMV_OK;
}

PMV_API int pmv_lk_track(pmv_ctx *ctx, const uint8_t *prev, const uint8_t *next,
                         int rows, int cols, int step, const float *prev_xy, int n,
                         int win_w, int win_h, int max_level, int max_count, double eps,
                         int flags, double min_eig_thr, float *next_xy, uint8_t *status, float *err)
{
    return pmv_lk_track_batched(ctx, prev, next, 1, (size_t)rows * step, rows, cols, step, prev_xy, n,
                                win_w, win_h, max_level, max_count, eps, flags, min_eig_thr,
                                next_xy, status, err);
}

}  // extern "C"
