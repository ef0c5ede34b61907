// corners.cu -- K4 min-eigenvalue response (GFTT flavour), K5 the reference's own Shi-Tomasi
// response (fp64), K6 corner selection (threshold, 3x3 NMS, sort, greedy min-distance, top-N).
//
// Replaces
//   cv::goodFeaturesToTrack(bw, corners, max, 0.01, 5, Mat(), 3, 3, false, 0.04)
//       reference OpenCVGoodFeatureExtractor.cpp:7           (SURVEY Appx B.1, B.2)
//   ShiTomasiFeatureExtractor::extractFeatures / computeShiTomasiResponse
//       reference ShiTomasiFeatureExtractor.cpp:5-75 + Frame.cpp:58-86,119-138   (Appx B.3)
//
// K4 arithmetic: the Sobel responses are small integers (|s| <= 1020), so the 3x3 box sums of
// their products are EXACT in int32; one fp32 scaling + OpenCV's closed form follows.  This is
// the same quantity OpenCV computes through rounded fp32 products (agreement ~1e-6 of the map
// maximum, inside the 1e-5 tie tolerance) at a third of the instruction count -- the kernel is
// meant to be HBM bound (1 B/px in, 4 B/px out).
// K5 arithmetic: gradients are half-integers, products quarter-integers -> box sums exact in
// int32, then the reference's fp64 expression order with explicit round-to-nearest intrinsics.
#include "common.cuh"
#include "sort.cuh"

namespace {

// ------------------------------------------------------------------ image tile staging ---
constexpr int CT_W = 64;   // output tile width  (one thread column each)
constexpr int CT_H = 32;   // output tile height (4 thread rows x 8 outputs)
constexpr int CT_R = 8;
constexpr int CS_W = CT_W + 4, CS_H = CT_H + 4;   // staged u8 tile (halo 2)
constexpr int CS_P = 72;                           // staged pitch (bytes)
constexpr int CD_W = CT_W + 2, CD_H = CT_H + 2;   // derivative tile (halo 1)

struct ImgView {            // uploaded patch of the parent image
    const uint8_t *ptr;     // device pointer to patch pixel (0,0)
    int pitch;
    int ox, oy;             // parent coordinates of patch pixel (0,0)
    int full_rows, full_cols;  // parent size (reflect-101 happens at the PARENT's edges)
    int rx, ry, rw, rh;     // ROI in parent coordinates
    int word_ok = 1;        // ROI column 0 sits on a 4-byte boundary of `ptr` rows (upload_roi guarantees it; a view of a
                            // resident image at an arbitrary ROI does not) -> aligned word loads allowed
};

constexpr int CS_X0 = 4;   // staged column of ROI column tx0 (the tile starts on a 4-byte boundary at tx0 - 4)

// stage ROI rows [ty0-2, ty0+CT_H+2) x ROI columns [tx0-4, tx0+CT_W+4) into s (u8, pitch CS_P = 72 B).
// Interior tiles (no parent edge inside the halo) are copied with aligned 32-bit loads -- upload_roi puts ROI
// column 0 on a 16-byte boundary; edge tiles fall back to byte gathers with reflect-101 at the PARENT's edges.
__device__ __forceinline__ void stage_tile(uint8_t (*s)[CS_P], const ImgView &v, int tx0, int ty0, int tid)
{
    const bool interior = (v.rx + tx0 - 2 >= 0) && (v.rx + tx0 + CT_W + 2 <= v.full_cols) &&
                          (v.ry + ty0 - 2 >= 0) && (v.ry + ty0 + CT_H + 2 <= v.full_rows);
    if (interior && v.word_ok) {
        const uint8_t *g = v.ptr + (ptrdiff_t)(v.ry + ty0 - 2 - v.oy) * v.pitch + (v.rx + tx0 - CS_X0 - v.ox);
        for (int i = tid; i < CS_H * (CS_P / 4); i += 256) {
            const int r = i / (CS_P / 4), wc = i - r * (CS_P / 4);
            reinterpret_cast<uint32_t *>(&s[r][0])[wc] = __ldg(reinterpret_cast<const uint32_t *>(g + (ptrdiff_t)r * v.pitch) + wc);
        }
    } else {
        for (int i = tid; i < CS_H * CS_W; i += 256) {
            int r = i / CS_W, c = i - r * CS_W;
            int gy = reflect101(v.ry + ty0 - 2 + r, v.full_rows) - v.oy;
            int gx = reflect101(v.rx + tx0 - 2 + c, v.full_cols) - v.ox;
            s[r][c + CS_X0 - 2] = __ldg(v.ptr + (ptrdiff_t)gy * v.pitch + gx);
        }
    }
}

__device__ __forceinline__ float block_max_f(float v, float *s_red)
{
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = threadIdx.x < 8 ? s_red[threadIdx.x] : 0.f;
        for (int o = 4; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    }
    return v;  // valid in thread 0
}

// Closed form of cornerMinEigenVal from the exact integer box sums of the Sobel products:
// eig = k * (0.5 (Sxx + Syy) - sqrt(0.25 (Sxx - Syy)^2 + Sxy^2)),  k = (1/3060)^2.
// Sxx +- Syy are formed in integers (no cancellation error), one approximate sqrt (<= 2 ulp).
__device__ __forceinline__ float mineig_from_sums(int sxx, int sxy, int syy)
{
    const float kf = (float)((1.0 / 3060.0) * (1.0 / 3060.0));
    const float Pf = (float)(sxx + syy), Qf = (float)(sxx - syy), Bf = (float)sxy;
    const float t = fmaf(0.25f * Qf, Qf, Bf * Bf);
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    return kf * fmaf(0.5f, Pf, -r);
}

// the same closed form from box sums held as exact fp32 integers (Sxx + Syy rounds once, like the int -> float
// conversion of the integer sum does)
__device__ __forceinline__ float mineig_from_fsums(float sxx, float sxy, float syy)
{
    const float kf = (float)((1.0 / 3060.0) * (1.0 / 3060.0));
    const float Pf = __fadd_rn(sxx, syy), Qf = __fsub_rn(sxx, syy);
    const float t = fmaf(0.25f * Qf, Qf, sxy * sxy);
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    return kf * fmaf(0.5f, Pf, -r);
}

// Regions of the register-resident fast kernel: a warp owns 120 columns x R rows (R = 16, 32 or 64, picked per launch).
// Regions start at column 4 / row 2 and
// tile the view; the LAST region of a row / column of regions is shifted back so that it ends inside the view (it
// overlaps its neighbour and writes the same values again).  What is left is a rim of 2 rows at the top and bottom,
// 4 columns on the left and 4..7 on the right: mineig_rim_kernel.  Small or unaligned views take the tile kernel.
constexpr int MF_W = 120;
struct FastGrid {
    int ok, R, nx, ny, xlast, ylast, xend;   // regions nx x ny of 120 x R outputs; output columns [4, xend), rows [2, rh - 2)
};
__host__ __device__ __forceinline__ FastGrid mineig_fast_grid(int rw, int rh, int word_ok, int R)
{
    FastGrid g;
    g.R = R;
    g.ok = word_ok && rw >= MF_W + 8 && rh >= R + 4;
    g.xlast = (rw - MF_W - 4) & ~3; g.ylast = rh - R - 2;
    g.nx = g.ok ? (g.xlast - 4 + MF_W - 1) / MF_W + 1 : 0;
    g.ny = g.ok ? (g.ylast - 2 + R - 1) / R + 1 : 0;
    g.xend = g.xlast + MF_W;
    return g;
}
// Region height: every region reads R + 4 input rows, so taller regions amortise the halo (64: 6 % extra rows, 16: 25 %),
// as long as there are enough regions to fill the GPU four times over (24 warps per SM): on 8 x 4K, 64-row regions
// (2.4 waves) lose more to the tail than they save (measured 104 vs 100 us).
inline int mineig_pick_rows(int rw, int rh, int batch, int sm_count)
{
    static const int forced = getenv("PMV_MINEIG_ROWS") ? atoi(getenv("PMV_MINEIG_ROWS")) : 0;
    if (forced >= 4 && forced % 4 == 0 && rh >= forced + 4) return forced;
    for (int R = 64; R > 16; R >>= 1) {
        if (rh < R + 4) continue;
        const FastGrid g = mineig_fast_grid(rw, rh, 1, R);
        if ((long long)g.nx * g.ny * batch >= 4LL * 24 * sm_count) return R;
    }
    return 16;
}

// byte k of w as the float 2^23 + b (one PRMT): differences of two such values are the exact pixel differences
__device__ __forceinline__ float biased_byte(unsigned w, unsigned sel)
{
    return __uint_as_float(__byte_perm(w, 0x4B000000u, sel));
}

// K4 (interior): lane L holds the 4 pixels at columns x0 - 4 + 4L of the current input row (one aligned 32-bit load),
// the two missing neighbours come from the adjacent lanes; Sobel, products, the 3x3 box sum and the closed form stay in
// registers while the warp walks down R + 4 input rows.  Lanes 1..30 write 4 responses each (16-byte stores when the
// map pitch allows).  HBM traffic: 1 B/px in, 4 B/px out.
// All of it is fp32 arithmetic on EXACT integers (|Sobel| <= 1020, products <= 2^20, box sums <= 9.4e6 < 2^24), so the
// values are those of the integer formulation bit for bit while the adds run on the FMA pipes instead of the
// half-rate integer ALU (which bounded the integer version: 22 ALU instructions per pixel), and products fold into
// the horizontal sums as FFMA.  Vertical 3-sums share the pair h(t-1) + h(t) between two consecutive output rows.
// State a lane carries from input row to input row (all exact small integers held in fp32).
struct MfState {
    float fb1[6], fb2[6];                                      // biased pixels of the previous row / the row before
    float dx1[4], dx2[4];                                      // horizontal differences of those rows
    float Hxx[2][4], Hxy[2][4], Hyy[2][4];                     // horizontal 3-sums of the two previous gradient rows
    float Pxx[4], Pxy[4], Pyy[4];                              // h(t-1) + h(t) of the last even step
};

// One input row.  STAGE 0: rows 0, 1 of a region (pixels only); 1: rows 2, 3 (gradient row, no output yet);
// 2: even output step (pair = h(t-1) + h(t), out = h(t-2) + pair); 3: odd output step (out = pair + h(t)).
template <int STAGE>
__device__ __forceinline__ void mf_row(MfState &S, unsigned w, float *__restrict__ o, bool writer, bool vec, float &vmax)
{
    const unsigned wl = __shfl_up_sync(0xffffffffu, w, 1), wr = __shfl_down_sync(0xffffffffu, w, 1);
    float fb[6], dx[4];
    fb[0] = biased_byte(wl, 0x7543); fb[1] = biased_byte(w, 0x7540); fb[2] = biased_byte(w, 0x7541);
    fb[3] = biased_byte(w, 0x7542); fb[4] = biased_byte(w, 0x7543); fb[5] = biased_byte(wr, 0x7540);
#pragma unroll
    for (int j = 0; j < 4; j++) dx[j] = fb[j + 2] - fb[j];
    if (STAGE >= 1) {
        float dv[6], sx[4], sy[4];
#pragma unroll
        for (int k = 0; k < 6; k++) dv[k] = fb[k] - S.fb2[k];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            sx[j] = fmaf(2.f, S.dx1[j], S.dx2[j]) + dx[j];
            sy[j] = fmaf(2.f, dv[j + 1], dv[j]) + dv[j + 2];
        }
        // horizontal 3-sums of the products; the outer terms come from the neighbouring lanes
        const float xx0 = sx[0] * sx[0], xy0 = sx[0] * sy[0], yy0 = sy[0] * sy[0];
        const float xx3 = sx[3] * sx[3], xy3 = sx[3] * sy[3], yy3 = sy[3] * sy[3];
        const float lxx = __shfl_up_sync(0xffffffffu, xx3, 1), lxy = __shfl_up_sync(0xffffffffu, xy3, 1),
                    lyy = __shfl_up_sync(0xffffffffu, yy3, 1);
        const float rxx = __shfl_down_sync(0xffffffffu, xx0, 1), rxy = __shfl_down_sync(0xffffffffu, xy0, 1),
                    ryy = __shfl_down_sync(0xffffffffu, yy0, 1);
        float hxx[4], hxy[4], hyy[4];
        {
            const float a = fmaf(sx[1], sx[1], xx0), b = fmaf(sx[2], sx[2], xx3);
            hxx[0] = lxx + a; hxx[1] = fmaf(sx[2], sx[2], a); hxx[2] = fmaf(sx[1], sx[1], b); hxx[3] = b + rxx;
        }
        {
            const float a = fmaf(sx[1], sy[1], xy0), b = fmaf(sx[2], sy[2], xy3);
            hxy[0] = lxy + a; hxy[1] = fmaf(sx[2], sy[2], a); hxy[2] = fmaf(sx[1], sy[1], b); hxy[3] = b + rxy;
        }
        {
            const float a = fmaf(sy[1], sy[1], yy0), b = fmaf(sy[2], sy[2], yy3);
            hyy[0] = lyy + a; hyy[1] = fmaf(sy[2], sy[2], a); hyy[2] = fmaf(sy[1], sy[1], b); hyy[3] = b + ryy;
        }
        if (STAGE >= 2) {
            float e[4];
            if (STAGE == 2) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    S.Pxx[j] = S.Hxx[1][j] + hxx[j]; S.Pxy[j] = S.Hxy[1][j] + hxy[j]; S.Pyy[j] = S.Hyy[1][j] + hyy[j];
                    e[j] = mineig_from_fsums(S.Hxx[0][j] + S.Pxx[j], S.Hxy[0][j] + S.Pxy[j], S.Hyy[0][j] + S.Pyy[j]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) e[j] = mineig_from_fsums(S.Pxx[j] + hxx[j], S.Pxy[j] + hxy[j], S.Pyy[j] + hyy[j]);
            }
            if (writer) {
                if (vec) *reinterpret_cast<float4 *>(o) = make_float4(e[0], e[1], e[2], e[3]);
                else { o[0] = e[0]; o[1] = e[1]; o[2] = e[2]; o[3] = e[3]; }
                vmax = fmaxf(vmax, fmaxf(fmaxf(e[0], e[1]), fmaxf(e[2], e[3])));
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            S.Hxx[0][j] = S.Hxx[1][j]; S.Hxy[0][j] = S.Hxy[1][j]; S.Hyy[0][j] = S.Hyy[1][j];
            S.Hxx[1][j] = hxx[j]; S.Hxy[1][j] = hxy[j]; S.Hyy[1][j] = hyy[j];
        }
    }
#pragma unroll
    for (int k = 0; k < 6; k++) { S.fb2[k] = S.fb1[k]; S.fb1[k] = fb[k]; }
#pragma unroll
    for (int j = 0; j < 4; j++) { S.dx2[j] = S.dx1[j]; S.dx1[j] = dx[j]; }
}

// The row loop is a ROLLED loop of four rows per trip (the state rotates with period two, so a four-row body needs no
// register moves) with the next four input rows requested at the top of each trip: the fully unrolled form of this
// kernel (R + 4 rows, ~18 k instructions) spent 0.85 stall cycles per issue waiting for instruction fetch.
__global__ void __launch_bounds__(256, 3)
mineig_fast_kernel(ImgView v, float *__restrict__ eig, int *__restrict__ max_bits, size_t vstride, size_t estride, int R)
{
    v.ptr += (size_t)blockIdx.z * vstride; eig += (size_t)blockIdx.z * estride; max_bits += blockIdx.z;   // image of a batch
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const FastGrid fg = mineig_fast_grid(v.rw, v.rh, v.word_ok, R);
    const int ky = blockIdx.y * 8 + warp;
    if (!fg.ok || ky >= fg.ny) return;                                      // warp-uniform
    const int x0 = min(4 + (int)blockIdx.x * MF_W, fg.xlast), y0 = min(2 + ky * R, fg.ylast);
    const int cx = x0 - 4 + 4 * lane;                          // ROI column of byte 0 of this lane's word
    const uint8_t *g = v.ptr + (ptrdiff_t)(v.ry + y0 - 2 - v.oy) * v.pitch + (v.rx + cx - v.ox);
    const bool writer = lane >= 1 && lane <= 30;
    const bool vec = (v.rw & 3) == 0;
    float *o = eig + (size_t)y0 * v.rw + cx;                   // output row y0 belongs to input row it = 4
    MfState S;
    float vmax = 0.f;
    unsigned cur[4], nxt[4];
#pragma unroll
    for (int k = 0; k < 4; k++) cur[k] = __ldg(reinterpret_cast<const unsigned *>(g + (ptrdiff_t)k * v.pitch));
#pragma unroll
    for (int k = 0; k < 4; k++) nxt[k] = __ldg(reinterpret_cast<const unsigned *>(g + (ptrdiff_t)(4 + k) * v.pitch));
    mf_row<0>(S, cur[0], o, writer, vec, vmax);
    mf_row<0>(S, cur[1], o, writer, vec, vmax);
    mf_row<1>(S, cur[2], o, writer, vec, vmax);
    mf_row<1>(S, cur[3], o, writer, vec, vmax);
#pragma unroll 1
    for (int it = 4; it < R + 4; it += 4) {
#pragma unroll
        for (int k = 0; k < 4; k++) cur[k] = nxt[k];
        if (it + 4 < R + 4) {
#pragma unroll
            for (int k = 0; k < 4; k++) nxt[k] = __ldg(reinterpret_cast<const unsigned *>(g + (ptrdiff_t)(it + 4 + k) * v.pitch));
        }
        mf_row<2>(S, cur[0], o, writer, vec, vmax);
        mf_row<3>(S, cur[1], o + v.rw, writer, vec, vmax);
        mf_row<2>(S, cur[2], o + 2 * (size_t)v.rw, writer, vec, vmax);
        mf_row<3>(S, cur[3], o + 3 * (size_t)v.rw, writer, vec, vmax);
        o += 4 * (size_t)v.rw;
    }
    for (int o2 = 16; o2; o2 >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o2));
    if (lane == 0 && vmax > 0.f) atomicMax(max_bits, __float_as_int(vmax));
}

// K4: eig(y,x) over the ROI + global max (atomicMax on the bits of a non-negative float)
__global__ void __launch_bounds__(256)
mineig_kernel(ImgView v, float *__restrict__ eig, int *__restrict__ max_bits, size_t vstride, size_t estride)
{
    v.ptr += (size_t)blockIdx.z * vstride; eig += (size_t)blockIdx.z * estride; max_bits += blockIdx.z;
    __shared__ __align__(16) uint8_t s_px[CS_H][CS_P];
    __shared__ int s_d[CD_H][CD_W];   // Sobel (sx | sy << 16) at ROI coords (ty0-1+r, tx0-1+c)
    __shared__ float s_red[8];
    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * CT_W, ty0 = blockIdx.y * CT_H;
    stage_tile(s_px, v, tx0, ty0, tid);
    __syncthreads();
    // Sobel at halo-1 positions; positions outside the ROI take the value of their reflect-101
    // image INSIDE the ROI (the covariance image is isolated, B.1)
    for (int i = tid; i < CD_H * CD_W; i += 256) {
        int r = i / CD_W, c = i - r * CD_W;
        int y = reflect101(ty0 - 1 + r, v.rh), x = reflect101(tx0 - 1 + c, v.rw);
        int sr = y - (ty0 - 2), sc = x - tx0 + CS_X0;   // staged indices of the centre
        int packed = 0;
        if (sr >= 1 && sr < CS_H - 1 && sc >= CS_X0 - 1 && sc <= CS_X0 + CT_W) {
            const uint8_t *p = &s_px[sr][sc];
            int a0 = p[-CS_P - 1], a1 = p[-CS_P], a2 = p[-CS_P + 1];
            int b0 = p[-1], b2 = p[1];
            int c0 = p[CS_P - 1], c1 = p[CS_P], c2 = p[CS_P + 1];
            int sx = (a2 - a0) + 2 * (b2 - b0) + (c2 - c0);
            int sy = (c0 - a0) + 2 * (c1 - a1) + (c2 - a2);
            packed = (sx & 0xffff) | (int)((unsigned)sy << 16);
        }
        s_d[r][c] = packed;
    }
    __syncthreads();
    const int tx = tid & 63, tyb = (tid >> 6) * CT_R;
    const int x = tx0 + tx;
    // sliding 3-row window of horizontal 3-sums of products
    int hxx[3], hxy[3], hyy[3];
    auto hsum = [&](int dr, int &oxx, int &oxy, int &oyy) {
        oxx = oxy = oyy = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            int d = s_d[dr][tx + k];
            int sx = (int)(short)d, sy = d >> 16;
            oxx += sx * sx; oxy += sx * sy; oyy += sy * sy;
        }
    };
    hsum(tyb + 0, hxx[0], hxy[0], hyy[0]);
    hsum(tyb + 1, hxx[1], hxy[1], hyy[1]);
    float vmax = 0.f;
#pragma unroll
    for (int r = 0; r < CT_R; r++) {
        hsum(tyb + r + 2, hxx[(r + 2) % 3], hxy[(r + 2) % 3], hyy[(r + 2) % 3]);
        const int y = ty0 + tyb + r;
        if (x < v.rw && y < v.rh) {
            const float e = mineig_from_sums(hxx[0] + hxx[1] + hxx[2], hxy[0] + hxy[1] + hxy[2], hyy[0] + hyy[1] + hyy[2]);
            eig[(size_t)y * v.rw + x] = e;
            vmax = fmaxf(vmax, e);
        }
    }
    vmax = block_max_f(vmax, s_red);
    if (tid == 0 && vmax > 0.f) atomicMax(max_bits, __float_as_int(vmax));
}

// reflect-101 for an overshoot of at most len - 1 (no loop)
__device__ __forceinline__ int reflect101_near(int p, int len)
{
    p = p < 0 ? -p : p;
    return p >= len ? 2 * (len - 1) - p : p;
}
// 3-tap box weights at position q of an isolated image of `len` samples with reflect-101 borders
__device__ __forceinline__ void rim_weights(int q, int len, int w[3])
{
    w[0] = w[1] = w[2] = 1;
    if (q == 0) { w[0] = 0; w[2] = 2; }
    if (q == len - 1) { w[2] = 0; w[0] = w[0] ? 2 : 0; }      // len >= 3 here, so both cannot apply
}

// K4 (rim of a view whose interior the register-resident kernel covers): one thread per pixel, everything from global
// memory.  Sobel taps read the PARENT (reflect-101 at its edges only); gradient positions outside the view take the
// value of their reflect-101 image inside it (the covariance image is isolated) -- the tile kernel's rules.
__global__ void __launch_bounds__(256)
mineig_rim_kernel(ImgView v, float *__restrict__ eig, int *__restrict__ max_bits, size_t vstride, size_t estride, int xend)
{
    v.ptr += (size_t)blockIdx.z * vstride; eig += (size_t)blockIdx.z * estride; max_bits += blockIdx.z;
    const int wr = v.rw - xend, side = 4 + wr;                 // rim columns per interior row: 4 left + wr right
    const int n_tb = 4 * v.rw, n = n_tb + (v.rh - 4) * side;
    const int i = blockIdx.x * 256 + threadIdx.x;
    float e = 0.f;
    if (i < n) {
        int x, y;
        if (i < n_tb) { const int r = i / v.rw; x = i - r * v.rw; y = r < 2 ? r : v.rh - 4 + r; }
        else { const int k = i - n_tb, r = k / side, c = k - r * side; y = 2 + r; x = c < 4 ? c : xend + c - 4; }
        // the 5 x 5 parent neighbourhood once (25 independent loads), then the nine Sobel responses from registers
        int p[5][5];
#pragma unroll
        for (int a = 0; a < 5; a++) {
            const int py = reflect101_near(v.ry + y - 2 + a, v.full_rows) - v.oy;
#pragma unroll
            for (int b = 0; b < 5; b++) {
                const int px = reflect101_near(v.rx + x - 2 + b, v.full_cols) - v.ox;
                p[a][b] = __ldg(v.ptr + (ptrdiff_t)py * v.pitch + px);
            }
        }
        // gradient positions outside the view fold back onto their reflect-101 image inside it: weights (0,1,2) / (2,1,0)
        // on the view's first / last row or column, (1,1,1) elsewhere
        int wr[3], wc[3];
        rim_weights(y, v.rh, wr); rim_weights(x, v.rw, wc);
        int sxx = 0, sxy = 0, syy = 0;
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) {
                const int sx = (p[a][b + 2] - p[a][b]) + 2 * (p[a + 1][b + 2] - p[a + 1][b]) + (p[a + 2][b + 2] - p[a + 2][b]);
                const int sy = (p[a + 2][b] - p[a][b]) + 2 * (p[a + 2][b + 1] - p[a][b + 1]) + (p[a + 2][b + 2] - p[a][b + 2]);
                const int wgt = wr[a] * wc[b];
                sxx += wgt * sx * sx; sxy += wgt * sx * sy; syy += wgt * sy * sy;
            }
        e = mineig_from_sums(sxx, sxy, syy);
        eig[(size_t)y * v.rw + x] = e;
    }
    for (int o = 16; o; o >>= 1) e = fmaxf(e, __shfl_xor_sync(0xffffffffu, e, o));
    if ((threadIdx.x & 31) == 0 && e > 0.f) atomicMax(max_bits, __float_as_int(e));
}

// K6a: threshold + 3x3 NMS + unordered compaction of (response bits, index) records.  A CTA sweeps a
// 32 x 64 pixel tile and gathers its candidates in shared memory, so the global counter sees one atomic per
// tile (an atomic per candidate on ONE address serialised in L2: 138 us for ~60 k candidates of a 4K frame).
constexpr int GC_ROWS = 64;

__global__ void __launch_bounds__(256)
gftt_candidates_kernel(const float *__restrict__ eig, int rows, int cols, const int *__restrict__ max_bits,
                       double quality, Rec128 *__restrict__ out, int *__restrict__ count, int cap, int *__restrict__ hist,
                       int *__restrict__ hist_overflow)
{
    __shared__ Rec128 s_rec[32 * GC_ROWS];   // every pixel of the tile may qualify (plateaus)
    __shared__ int s_n, s_base;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const float thr = (float)((double)__int_as_float(*max_bits) * quality);
    // the eight responses of this thread's column are requested together (few of them pass the threshold: the kernel is one
    // streaming read of the map, and eight dependent load -> test -> branch rounds left it latency-bound)
    float vals[GC_ROWS / 8];
#pragma unroll
    for (int k = 0; k < GC_ROWS / 8; k++) {
        const int y = blockIdx.y * GC_ROWS + 8 * k + (threadIdx.x >> 5);
        const bool in = x >= 1 && y >= 1 && x < cols - 1 && y < rows - 1;
        vals[k] = in ? __ldg(eig + (size_t)y * cols + x) : -1.f;
    }
#pragma unroll
    for (int k = 0; k < GC_ROWS / 8; k++) {
        const int y = blockIdx.y * GC_ROWS + 8 * k + (threadIdx.x >> 5);
        const float val = vals[k];
        if (!(val > thr)) continue;       // also skips the positions outside the interior (-1; thr >= 0)
        const float *p = eig + (size_t)y * cols + x;
        float m = fmaxf(fmaxf(p[-cols - 1], p[-cols]), fmaxf(p[-cols + 1], p[-1]));
        m = fmaxf(m, fmaxf(fmaxf(p[1], p[cols - 1]), fmaxf(p[cols], p[cols + 1])));
        if (val < m) continue;  // val == dilate(thresholded) <=> val >= every neighbour
        const int slot = atomicAdd(&s_n, 1);
        s_rec[slot] = Rec128{(unsigned long long)__float_as_uint(val), (unsigned long long)(y * cols + x)};
        // response histogram for the bucket sort (sort.cuh)
        if (hist) {
            const int b = bs_bucket(__float_as_uint(val), (unsigned)*max_bits);
            if (b >= 0 && b < BS_BINS) atomicAdd(&hist[b], 1);
            else *hist_overflow = 0x7fffffff;
        }
    }
    __syncthreads();
    const int n = s_n;
    if (n == 0) return;
    if (threadIdx.x == 0) s_base = atomicAdd(count, n);
    __syncthreads();
    const int base = s_base;
    for (int i = threadIdx.x; i < n; i += 256)
        if (base + i < cap) out[base + i] = s_rec[i];
}

__global__ void __launch_bounds__(256)
rank_scatter_kernel(const Rec128 *__restrict__ recs, int n, int *__restrict__ rankmap)
{
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) rankmap[(int)recs[i].lo] = i;
}

// K6a': for every candidate the ranks of the candidates of SMALLER rank (stronger response) closer than
// min_dist -- the only ones whose acceptance can reject it.  Done by the whole GPU once, so the single-CTA
// greedy pass below checks a handful of status bytes per candidate instead of scanning (2R+1)^2 pixels of the
// rank map in every fixed-point iteration (0.63 ms -> the rank-map scan was 90 % of the selection at 4K).
// nbc[i] = 255: more than GF_NBMAX such neighbours (plateaus), the greedy pass scans the window itself.
constexpr int GF_NBMAX = 16;

__global__ void __launch_bounds__(256)
gftt_neighbors_kernel(const Rec128 *__restrict__ recs, int n, int rows, int cols, const int *__restrict__ rankmap,
                      float min_dist, int *__restrict__ nb, unsigned char *__restrict__ nbc)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float md2 = min_dist * min_dist;
    int R = (int)min_dist;
    if ((float)R >= min_dist) R -= 1;
    if (R < 0) R = 0;
    int cnt = 0;
    if (min_dist >= 1.f) {
        const int idx = (int)recs[i].lo;
        const int y = idx / cols, x = idx - y * cols;
        for (int dy = -R; dy <= R; dy++) {
            const int yy = y + dy;
            if (yy < 0 || yy >= rows) continue;
            for (int dx = -R; dx <= R; dx++) {
                const int xx = x + dx;
                if (xx < 0 || xx >= cols) continue;
                if ((float)(dx * dx + dy * dy) >= md2) continue;
                const int rk = rankmap[yy * cols + xx];
                if (rk < 0 || rk >= i) continue;
                if (cnt < GF_NBMAX) nb[(size_t)i * GF_NBMAX + cnt] = rk;
                cnt++;
            }
        }
    }
    nbc[i] = cnt > GF_NBMAX ? (unsigned char)255 : (unsigned char)cnt;
}

// K6b: greedy min-distance selection == the sequential loop of goodFeaturesToTrack, evaluated in
// rank-ordered chunks of 1024 by ONE CTA: candidate i is accepted iff no ACCEPTED candidate of
// smaller rank lies within min_dist; inside a chunk the decision is iterated to its fixed point.
__global__ void __launch_bounds__(1024)
gftt_select_kernel(const Rec128 *__restrict__ recs, int n, int rows, int cols, const int *__restrict__ rankmap,
                   const int *__restrict__ nb, const unsigned char *__restrict__ nbc,
                   unsigned char *status /* n, zeroed */, float min_dist, int max_corners,
                   float *__restrict__ out_xy, float *__restrict__ out_score, int *__restrict__ out_n)
{
    __shared__ int s_warp[32];
    __shared__ int s_total, s_chunk;
    const int tid = threadIdx.x;
    if (tid == 0) s_total = 0;
    __syncthreads();
    const float md2 = min_dist * min_dist;
    int R = (int)min_dist;
    if ((float)R >= min_dist) R -= 1;  // largest integer offset with d*d < min_dist^2
    if (R < 0) R = 0;
    volatile unsigned char *vst = status;
    const int limit = max_corners > 0 ? max_corners : n;
    for (int base = 0; base < n; base += 1024) {
        if (s_total >= limit) break;
        const int i = base + tid;
        const bool active = i < n;
        int x = 0, y = 0;
        float val = 0.f;
        if (active) {
            Rec128 r = recs[i];
            int idx = (int)r.lo;
            y = idx / cols; x = idx - y * cols;
            val = __uint_as_float((unsigned)r.hi);
        }
        int st = 0;
        const int ncnt = active ? (int)nbc[i] : 0;
        int nbr[GF_NBMAX];
        if (ncnt != 255) {
#pragma unroll
            for (int k = 0; k < GF_NBMAX; k++) nbr[k] = k < ncnt ? nb[(size_t)i * GF_NBMAX + k] : -1;
        }
        while (true) {
            if (active && st == 0) {
                bool pending = false, rejected = false;
                if (ncnt != 255) {
#pragma unroll
                    for (int k = 0; k < GF_NBMAX; k++) {
                        if (k < ncnt) {
                            const unsigned char sk = vst[nbr[k]];
                            if (sk == 1) rejected = true;
                            if (sk == 0) pending = true;
                        }
                    }
                } else if (min_dist >= 1.f) {
                    for (int dy = -R; dy <= R && !rejected; dy++) {
                        int yy = y + dy;
                        if (yy < 0 || yy >= rows) continue;
                        for (int dx = -R; dx <= R; dx++) {
                            int xx = x + dx;
                            if (xx < 0 || xx >= cols) continue;
                            if ((float)(dx * dx + dy * dy) >= md2) continue;
                            int rk = rankmap[yy * cols + xx];
                            if (rk < 0 || rk >= i) continue;
                            unsigned char s = vst[rk];
                            if (s == 1) { rejected = true; break; }
                            if (s == 0) pending = true;
                        }
                    }
                }
                if (rejected) st = 2; else if (!pending) st = 1;
                if (st) vst[i] = (unsigned char)st;
            }
            __threadfence_block();
            if (!__syncthreads_or(active && st == 0)) break;
        }
        // ordered output of the accepted candidates of this chunk
        const int acc = (active && st == 1) ? 1 : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, acc);
        const int lane = tid & 31, warp = tid >> 5;
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        if (warp == 0) {
            int v = s_warp[lane], incl = v;
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            s_warp[lane] = incl - v;  // exclusive prefix per warp
            if (lane == 31) s_chunk = incl;
        }
        __syncthreads();
        const int before = s_total;
        const int pos = before + s_warp[warp] + __popc(bal & ((1u << lane) - 1));
        if (acc && pos < limit) {
            out_xy[2 * pos] = (float)x;
            out_xy[2 * pos + 1] = (float)y;
            out_score[pos] = val;
        }
        __syncthreads();
        if (tid == 0) s_total = before + s_chunk;
        __syncthreads();
    }
    if (tid == 0) *out_n = s_total < limit ? s_total : limit;
}

// ------------------------------------------------------------------ K5 reference ShiTomasi
// min(lambda_1, lambda_2) of the blurred Harris matrix from the box sums of (2 gx)^2, (2 gx)(2 gy), (2 gy)^2 (exact
// integers): / 4 is exact, then the reference's fp64 expression in its own order, every operation rounded to nearest
// (ShiTomasiFeatureExtractor.cpp:62-71, Frame.cpp:119-138 blur = sum * (1/9))
__device__ __forceinline__ double shitomasi_from_sums(double sxx, double sxy, double syy)
{
    const double ninth = 1.0 / 9;
    const double Ixx = __dmul_rn(sxx * 0.25, ninth), Iyy = __dmul_rn(syy * 0.25, ninth), Ixy = __dmul_rn(sxy * 0.25, ninth);
    const double B = __dsub_rn(-Ixx, Iyy);
    const double C = __dsub_rn(__dmul_rn(Ixx, Iyy), __dmul_rn(Ixy, Ixy));
    const double disc = __dsub_rn(__dmul_rn(B, B), __dmul_rn(4.0, C));
    const double sq = __dsqrt_rn(disc);
    const double l1 = __dmul_rn(__dadd_rn(-B, sq), 0.5);
    const double l2 = __dmul_rn(__dsub_rn(-B, sq), 0.5);
    return (l2 < l1) ? l2 : l1;  // std::min(l1, l2)
}

// R(y,x) in fp64 + global max (bits of a non-negative double).  The view is isolated (fresh Mats in
// the reference): gradient zero on the rim, blur reflects at the view's own edges, last column 0.
__global__ void __launch_bounds__(256)
shitomasi_response_kernel(ImgView v, int signed_quirk, double *__restrict__ R,
                          unsigned long long *__restrict__ max_bits, size_t vstride, size_t rstride)
{
    v.ptr += (size_t)blockIdx.z * vstride; R += (size_t)blockIdx.z * rstride; max_bits += blockIdx.z;
    __shared__ __align__(16) uint8_t s_px[CS_H][CS_P];
    __shared__ int s_d[CD_H][CD_W];   // (2*gx) | (2*gy) << 16 at view coords (ty0-1+r, tx0-1+c)
    __shared__ double s_redd[8];
    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * CT_W, ty0 = blockIdx.y * CT_H;
    stage_tile(s_px, v, tx0, ty0, tid);
    __syncthreads();
    for (int i = tid; i < CD_H * CD_W; i += 256) {
        int r = i / CD_W, c = i - r * CD_W;
        int y = reflect101(ty0 - 1 + r, v.rh), x = reflect101(tx0 - 1 + c, v.rw);
        int sr = y - (ty0 - 2), sc = x - tx0 + CS_X0;
        int packed = 0;
        // Frame.cpp:63-84: interior pixels only, rim gradients stay zero
        if (y >= 1 && y < v.rh - 1 && x >= 1 && x < v.rw - 1 && sr >= 1 && sr < CS_H - 1 && sc >= CS_X0 - 1 && sc <= CS_X0 + CT_W) {
            const uint8_t *p = &s_px[sr][sc];
            int l = p[-1], rr = p[1], u = p[-CS_P], d = p[CS_P];
            if (signed_quirk) { l = (signed char)l; rr = (signed char)rr; u = (signed char)u; d = (signed char)d; }
            int gx2 = rr - l, gy2 = d - u;   // 2*gx, 2*gy (exact)
            packed = (gx2 & 0xffff) | (int)((unsigned)gy2 << 16);
        }
        s_d[r][c] = packed;
    }
    __syncthreads();
    const int tx = tid & 63, tyb = (tid >> 6) * CT_R;
    const int x = tx0 + tx;
    int hxx[3], hxy[3], hyy[3];
    auto hsum = [&](int dr, int &oxx, int &oxy, int &oyy) {
        oxx = oxy = oyy = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            int d = s_d[dr][tx + k];
            int gx = (int)(short)d, gy = d >> 16;
            oxx += gx * gx; oxy += gx * gy; oyy += gy * gy;
        }
    };
    hsum(tyb + 0, hxx[0], hxy[0], hyy[0]);
    hsum(tyb + 1, hxx[1], hxy[1], hyy[1]);
    double vmax = 0.0;
#pragma unroll
    for (int r = 0; r < CT_R; r++) {
        hsum(tyb + r + 2, hxx[(r + 2) % 3], hxy[(r + 2) % 3], hyy[(r + 2) % 3]);
        const int y = ty0 + tyb + r;
        if (x < v.rw && y < v.rh) {
            // ShiTomasiFeatureExtractor.cpp:58 skips the last column
            const double out = x < v.rw - 1 ? shitomasi_from_sums((double)(hxx[0] + hxx[1] + hxx[2]), (double)(hxy[0] + hxy[1] + hxy[2]),
                                                                   (double)(hyy[0] + hyy[1] + hyy[2])) : 0.0;
            R[(size_t)y * v.rw + x] = out;
            if (out > vmax) vmax = out;
        }
    }
    for (int o = 16; o; o >>= 1) vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if ((tid & 31) == 0) s_redd[tid >> 5] = vmax;
    __syncthreads();
    if (tid == 0) {
        for (int k = 1; k < 8; k++) vmax = fmax(vmax, s_redd[k]);
        if (vmax > 0.0) atomicMax(max_bits, (unsigned long long)__double_as_longlong(vmax));
    }
}

// K5 (interior), the register-resident scheme of mineig_fast_kernel for the reference's response: central differences
// 2 gx = right - left, 2 gy = down - up (Frame.cpp:63-84) as exact fp32 integers, products folded into the horizontal
// 3-sums, vertical 3-sums sharing a pair between two output rows; only the closed form runs in fp64.
struct SfState {
    float fb1[6], fb2[6];                                      // biased pixels of the previous row / the row before
    float Hxx[2][4], Hxy[2][4], Hyy[2][4];
    float Pxx[4], Pxy[4], Pyy[4];
};

template <int STAGE>
__device__ __forceinline__ void sf_row(SfState &S, unsigned w, unsigned flip, double *__restrict__ o, bool writer, bool vec, double &vmax)
{
    w ^= flip;                                                 // signed-char quirk: b ^ 0x80 = (signed char)b + 128
    const unsigned wl = __shfl_up_sync(0xffffffffu, w, 1), wr = __shfl_down_sync(0xffffffffu, w, 1);
    float fb[6];
    fb[0] = biased_byte(wl, 0x7543); fb[1] = biased_byte(w, 0x7540); fb[2] = biased_byte(w, 0x7541);
    fb[3] = biased_byte(w, 0x7542); fb[4] = biased_byte(w, 0x7543); fb[5] = biased_byte(wr, 0x7540);
    if (STAGE >= 1) {
        float gx[4], gy[4];                                    // gradient row = the previous input row
#pragma unroll
        for (int j = 0; j < 4; j++) { gx[j] = S.fb1[j + 2] - S.fb1[j]; gy[j] = fb[j + 1] - S.fb2[j + 1]; }
        const float xx0 = gx[0] * gx[0], xy0 = gx[0] * gy[0], yy0 = gy[0] * gy[0];
        const float xx3 = gx[3] * gx[3], xy3 = gx[3] * gy[3], yy3 = gy[3] * gy[3];
        const float lxx = __shfl_up_sync(0xffffffffu, xx3, 1), lxy = __shfl_up_sync(0xffffffffu, xy3, 1),
                    lyy = __shfl_up_sync(0xffffffffu, yy3, 1);
        const float rxx = __shfl_down_sync(0xffffffffu, xx0, 1), rxy = __shfl_down_sync(0xffffffffu, xy0, 1),
                    ryy = __shfl_down_sync(0xffffffffu, yy0, 1);
        float hxx[4], hxy[4], hyy[4];
        {
            const float a = fmaf(gx[1], gx[1], xx0), b = fmaf(gx[2], gx[2], xx3);
            hxx[0] = lxx + a; hxx[1] = fmaf(gx[2], gx[2], a); hxx[2] = fmaf(gx[1], gx[1], b); hxx[3] = b + rxx;
        }
        {
            const float a = fmaf(gx[1], gy[1], xy0), b = fmaf(gx[2], gy[2], xy3);
            hxy[0] = lxy + a; hxy[1] = fmaf(gx[2], gy[2], a); hxy[2] = fmaf(gx[1], gy[1], b); hxy[3] = b + rxy;
        }
        {
            const float a = fmaf(gy[1], gy[1], yy0), b = fmaf(gy[2], gy[2], yy3);
            hyy[0] = lyy + a; hyy[1] = fmaf(gy[2], gy[2], a); hyy[2] = fmaf(gy[1], gy[1], b); hyy[3] = b + ryy;
        }
        if (STAGE >= 2) {
            double e[4];
            if (STAGE == 2) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    S.Pxx[j] = S.Hxx[1][j] + hxx[j]; S.Pxy[j] = S.Hxy[1][j] + hxy[j]; S.Pyy[j] = S.Hyy[1][j] + hyy[j];
                    e[j] = shitomasi_from_sums((double)(S.Hxx[0][j] + S.Pxx[j]), (double)(S.Hxy[0][j] + S.Pxy[j]), (double)(S.Hyy[0][j] + S.Pyy[j]));
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++)
                    e[j] = shitomasi_from_sums((double)(S.Pxx[j] + hxx[j]), (double)(S.Pxy[j] + hxy[j]), (double)(S.Pyy[j] + hyy[j]));
            }
            if (writer) {
                if (vec) {
                    *reinterpret_cast<double2 *>(o) = make_double2(e[0], e[1]);
                    *reinterpret_cast<double2 *>(o + 2) = make_double2(e[2], e[3]);
                } else { o[0] = e[0]; o[1] = e[1]; o[2] = e[2]; o[3] = e[3]; }
                vmax = fmax(vmax, fmax(fmax(e[0], e[1]), fmax(e[2], e[3])));
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            S.Hxx[0][j] = S.Hxx[1][j]; S.Hxy[0][j] = S.Hxy[1][j]; S.Hyy[0][j] = S.Hyy[1][j];
            S.Hxx[1][j] = hxx[j]; S.Hxy[1][j] = hxy[j]; S.Hyy[1][j] = hyy[j];
        }
    }
#pragma unroll
    for (int k = 0; k < 6; k++) { S.fb2[k] = S.fb1[k]; S.fb1[k] = fb[k]; }
}

__global__ void __launch_bounds__(256, 3)
shitomasi_fast_kernel(ImgView v, int signed_quirk, double *__restrict__ Rm, unsigned long long *__restrict__ max_bits, size_t vstride,
                      size_t rstride, int R)
{
    v.ptr += (size_t)blockIdx.z * vstride; Rm += (size_t)blockIdx.z * rstride; max_bits += blockIdx.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const FastGrid fg = mineig_fast_grid(v.rw, v.rh, v.word_ok, R);
    const int ky = blockIdx.y * 8 + warp;
    if (!fg.ok || ky >= fg.ny) return;                                      // warp-uniform
    const int x0 = min(4 + (int)blockIdx.x * MF_W, fg.xlast), y0 = min(2 + ky * R, fg.ylast);
    const int cx = x0 - 4 + 4 * lane;
    const uint8_t *g = v.ptr + (ptrdiff_t)(v.ry + y0 - 2 - v.oy) * v.pitch + (v.rx + cx - v.ox);
    const bool writer = lane >= 1 && lane <= 30;
    const bool vec = (v.rw & 1) == 0;
    const unsigned flip = signed_quirk ? 0x80808080u : 0u;
    double *o = Rm + (size_t)y0 * v.rw + cx;
    SfState S;
    double vmax = 0.0;
    unsigned cur[4], nxt[4];
#pragma unroll
    for (int k = 0; k < 4; k++) cur[k] = __ldg(reinterpret_cast<const unsigned *>(g + (ptrdiff_t)k * v.pitch));
#pragma unroll
    for (int k = 0; k < 4; k++) nxt[k] = __ldg(reinterpret_cast<const unsigned *>(g + (ptrdiff_t)(4 + k) * v.pitch));
    sf_row<0>(S, cur[0], flip, o, writer, vec, vmax);
    sf_row<0>(S, cur[1], flip, o, writer, vec, vmax);
    sf_row<1>(S, cur[2], flip, o, writer, vec, vmax);
    sf_row<1>(S, cur[3], flip, o, writer, vec, vmax);
#pragma unroll 1
    for (int it = 4; it < R + 4; it += 4) {
#pragma unroll
        for (int k = 0; k < 4; k++) cur[k] = nxt[k];
        if (it + 4 < R + 4) {
#pragma unroll
            for (int k = 0; k < 4; k++) nxt[k] = __ldg(reinterpret_cast<const unsigned *>(g + (ptrdiff_t)(it + 4 + k) * v.pitch));
        }
        sf_row<2>(S, cur[0], flip, o, writer, vec, vmax);
        sf_row<3>(S, cur[1], flip, o + v.rw, writer, vec, vmax);
        sf_row<2>(S, cur[2], flip, o + 2 * (size_t)v.rw, writer, vec, vmax);
        sf_row<3>(S, cur[3], flip, o + 3 * (size_t)v.rw, writer, vec, vmax);
        o += 4 * (size_t)v.rw;
    }
    for (int o2 = 16; o2; o2 >>= 1) vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o2));
    if (lane == 0 && vmax > 0.0) atomicMax(max_bits, (unsigned long long)__double_as_longlong(vmax));
}

// K5 (rim of a view whose interior shitomasi_fast_kernel covers): one thread per pixel.  The view is isolated: gradients
// are zero on its outermost rows / columns, the blur reflects (101) at the view's own edges, the last column stays 0.
__global__ void __launch_bounds__(256)
shitomasi_rim_kernel(ImgView v, int signed_quirk, double *__restrict__ Rm, unsigned long long *__restrict__ max_bits, size_t vstride,
                     size_t rstride, int xend)
{
    v.ptr += (size_t)blockIdx.z * vstride; Rm += (size_t)blockIdx.z * rstride; max_bits += blockIdx.z;
    const int wr_ = v.rw - xend, side = 4 + wr_;
    const int n_tb = 4 * v.rw, n = n_tb + (v.rh - 4) * side;
    const int i = blockIdx.x * 256 + threadIdx.x;
    double e = 0.0;
    if (i < n) {
        int x, y;
        if (i < n_tb) { const int r = i / v.rw; x = i - r * v.rw; y = r < 2 ? r : v.rh - 4 + r; }
        else { const int k = i - n_tb, r = k / side, c = k - r * side; y = 2 + r; x = c < 4 ? c : xend + c - 4; }
        if (x < v.rw - 1) {
            int p[5][5];
#pragma unroll
            for (int a = 0; a < 5; a++) {
                const int py = min(max(y - 2 + a, 0), v.rh - 1) + v.ry - v.oy;     // taps of interior gradients never leave the view
#pragma unroll
                for (int b = 0; b < 5; b++) {
                    const int px = min(max(x - 2 + b, 0), v.rw - 1) + v.rx - v.ox;
                    const int q = __ldg(v.ptr + (ptrdiff_t)py * v.pitch + px);
                    p[a][b] = signed_quirk ? (int)(signed char)q : q;
                }
            }
            int wr[3], wc[3];
            rim_weights(y, v.rh, wr); rim_weights(x, v.rw, wc);
            int sxx = 0, sxy = 0, syy = 0;
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int b = 0; b < 3; b++) {
                    const int gy_ = y - 1 + a, gx_ = x - 1 + b;                     // gradient position (inside the view wherever its weight is not 0)
                    const bool inner = gy_ >= 1 && gy_ < v.rh - 1 && gx_ >= 1 && gx_ < v.rw - 1;   // Frame.cpp:63-84: rim gradients stay zero
                    const int gx2 = inner ? p[a + 1][b + 2] - p[a + 1][b] : 0, gy2 = inner ? p[a + 2][b + 1] - p[a][b + 1] : 0;
                    const int wgt = wr[a] * wc[b];
                    sxx += wgt * gx2 * gx2; sxy += wgt * gx2 * gy2; syy += wgt * gy2 * gy2;
                }
            e = shitomasi_from_sums((double)sxx, (double)sxy, (double)syy);
        }
        Rm[(size_t)y * v.rw + x] = e;
    }
    for (int o = 16; o; o >>= 1) e = fmax(e, __shfl_xor_sync(0xffffffffu, e, o));
    if ((threadIdx.x & 31) == 0 && e > 0.0) atomicMax(max_bits, (unsigned long long)__double_as_longlong(e));
}

// K6 (reference flavour): every pixel with R > rmax*quality, no NMS
__global__ void __launch_bounds__(256)
shitomasi_candidates_kernel(const double *__restrict__ R, int n, const unsigned long long *__restrict__ max_bits,
                            double quality, Rec128 *__restrict__ out, int *__restrict__ count, int cap)
{
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const double thr = __dmul_rn(__longlong_as_double((long long)*max_bits), quality);
    const double v = R[i];
    if (!(v > thr)) return;
    int slot = atomicAdd(count, 1);
    // ties (std::sort is unstable in the reference): lower raster index first
    if (slot < cap) out[slot] = Rec128{(unsigned long long)__double_as_longlong(v), ~(unsigned long long)i};
}

__global__ void __launch_bounds__(256)
shitomasi_emit_kernel(const Rec128 *__restrict__ recs, int n, int cols, int *__restrict__ col, int *__restrict__ row,
                      double *__restrict__ score)
{
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    int idx = (int)(~recs[i].lo);
    row[i] = idx / cols; col[i] = idx - (idx / cols) * cols;
    score[i] = __longlong_as_double((long long)recs[i].hi);
}

// ------------------------------------------------------------------ host-side helpers -----
// Upload the ROI plus a 2 px halo (clipped to the parent) of a host image into ctx->img[0].
int upload_roi(pmv_ctx *ctx, const uint8_t *base, int full_rows, int full_cols, int step,
               int rx, int ry, int rw, int rh, ImgView *v)
{
    int x0 = rx - 2 < 0 ? 0 : rx - 2, y0 = ry - 2 < 0 ? 0 : ry - 2;
    int x1 = rx + rw + 2 > full_cols ? full_cols : rx + rw + 2;
    int y1 = ry + rh + 2 > full_rows ? full_rows : ry + rh + 2;
    int pw = x1 - x0, ph = y1 - y0;
    const int lead = 16 - (rx - x0);                 // ROI column 0 lands on a 16-byte boundary (tile word loads)
    int pitch = align_up(lead + pw + 16, 128);
    cudaError_t e = ctx->img[0].reserve((size_t)pitch * ph + 256);
    if (e != cudaSuccess) return ctx->fail(PMV_ERR_NOMEM, "roi upload buffer", e);
    uint8_t *p0 = ctx->img[0].as<uint8_t>() + lead;
    PMV_CUDA_TRY(ctx, cudaMemcpy2DAsync(p0, pitch, base + (size_t)y0 * step + x0, step, pw, ph,
                                        cudaMemcpyHostToDevice, ctx->stream));
    *v = ImgView{p0, pitch, x0, y0, full_rows, full_cols, rx, ry, rw, rh};
    return PMV_OK;
}

int check_roi(pmv_ctx *ctx, const void *base, int full_rows, int full_cols, int step, int rx, int ry, int rw, int rh)
{
    if (!base || full_rows <= 0 || full_cols <= 0 || step < full_cols)
        return ctx->fail(PMV_ERR_INVALID, "corner detector: bad image argument");
    if (rw <= 0 || rh <= 0 || rx < 0 || ry < 0 || rx + rw > full_cols || ry + rh > full_rows)
        return ctx->fail(PMV_ERR_INVALID, "corner detector: ROI outside the image");
    return PMV_OK;
}

// fp64 response map of the reference extractor + its maximum (the caller zeroes d_max): interior by the register-resident
// kernel, rim by one thread per pixel on the side stream; small or unaligned views by the tile kernel
int run_shitomasi(pmv_ctx *ctx, const ImgView &v, int signed_quirk, double *d_R, unsigned long long *d_max, cudaStream_t s, int batch = 1,
                  size_t vstride = 0, size_t rstride = 0)
{
    const FastGrid fg = mineig_fast_grid(v.rw, v.rh, v.word_ok, mineig_pick_rows(v.rw, v.rh, batch, ctx->sm_count));
    if (fg.ok) {
        const bool fork = ctx->copy_stream != nullptr && ctx->copy_stream != s && ctx->ev[2] != nullptr && ctx->ev[3] != nullptr;
        cudaStream_t se = fork ? ctx->copy_stream : s;
        if (fork) {
            PMV_CUDA_TRY(ctx, cudaEventRecord(ctx->ev[2], s));
            PMV_CUDA_TRY(ctx, cudaStreamWaitEvent(se, ctx->ev[2], 0));
        }
        const int n_rim = 4 * v.rw + (v.rh - 4) * (4 + v.rw - fg.xend);
        shitomasi_rim_kernel<<<dim3((n_rim + 255) / 256, 1, batch), 256, 0, se>>>(v, signed_quirk, d_R, d_max, vstride, rstride, fg.xend);
        PMV_LAUNCH_CHECK(ctx, "shitomasi_rim_kernel");
        shitomasi_fast_kernel<<<dim3(fg.nx, (fg.ny + 7) / 8, batch), 256, 0, s>>>(v, signed_quirk, d_R, d_max, vstride, rstride, fg.R);
        PMV_LAUNCH_CHECK(ctx, "shitomasi_fast_kernel");
        if (fork) {
            PMV_CUDA_TRY(ctx, cudaEventRecord(ctx->ev[3], se));
            PMV_CUDA_TRY(ctx, cudaStreamWaitEvent(s, ctx->ev[3], 0));
        }
    } else {
        dim3 grid((v.rw + CT_W - 1) / CT_W, (v.rh + CT_H - 1) / CT_H, batch);
        shitomasi_response_kernel<<<grid, 256, 0, s>>>(v, signed_quirk, d_R, d_max, vstride, rstride);
        PMV_LAUNCH_CHECK(ctx, "shitomasi_response_kernel");
    }
    return PMV_OK;
}

int run_mineig(pmv_ctx *ctx, const ImgView &v, float *d_eig, int *d_max, cudaStream_t s, int batch = 1, size_t vstride = 0, size_t estride = 0)
{
    ProfScope ps(ctx, PMV_PHASE_RESPONSE, s);
    PMV_CUDA_TRY(ctx, cudaMemsetAsync(d_max, 0, sizeof(int) * batch, s));
    const FastGrid fg = mineig_fast_grid(v.rw, v.rh, v.word_ok, mineig_pick_rows(v.rw, v.rh, batch, ctx->sm_count));
    if (fg.ok) {
        // rim (2 rows / 4..7 columns, one thread per pixel) on the side stream while the register-resident kernel covers
        // the interior: disjoint pixels, both feed the atomic maximum
        const bool fork = ctx->copy_stream != nullptr && ctx->copy_stream != s && ctx->ev[2] != nullptr && ctx->ev[3] != nullptr;
        cudaStream_t se = fork ? ctx->copy_stream : s;
        if (fork) {
            PMV_CUDA_TRY(ctx, cudaEventRecord(ctx->ev[2], s));
            PMV_CUDA_TRY(ctx, cudaStreamWaitEvent(se, ctx->ev[2], 0));
        }
        const int n_rim = 4 * v.rw + (v.rh - 4) * (4 + v.rw - fg.xend);
        mineig_rim_kernel<<<dim3((n_rim + 255) / 256, 1, batch), 256, 0, se>>>(v, d_eig, d_max, vstride, estride, fg.xend);
        PMV_LAUNCH_CHECK(ctx, "mineig_rim_kernel");
        dim3 gridf(fg.nx, (fg.ny + 7) / 8, batch);
        mineig_fast_kernel<<<gridf, 256, 0, s>>>(v, d_eig, d_max, vstride, estride, fg.R);
        PMV_LAUNCH_CHECK(ctx, "mineig_fast_kernel");
        if (fork) {
            PMV_CUDA_TRY(ctx, cudaEventRecord(ctx->ev[3], se));
            PMV_CUDA_TRY(ctx, cudaStreamWaitEvent(s, ctx->ev[3], 0));
        }
    } else {
        dim3 grid((v.rw + CT_W - 1) / CT_W, (v.rh + CT_H - 1) / CT_H, batch);   // small / unaligned views: tile kernel
        mineig_kernel<<<grid, 256, 0, s>>>(v, d_eig, d_max, vstride, estride);
        PMV_LAUNCH_CHECK(ctx, "mineig_kernel");
    }
    return PMV_OK;
}

// goodFeaturesToTrack on a view that is already on the device; the corner list stays there (scratch[5]).
int gftt_run(pmv_ctx *ctx, const ImgView &v, int max_corners, double quality, double min_dist, float **out_xy, float **out_sc,
             int *n_out)
{
    cudaStream_t s = ctx->stream;
    int rc;
    const size_t npx = (size_t)v.rw * v.rh;
    const int cap = (int)npx;  // plateaus of equal responses can make every pixel a candidate
    const int cap2 = sort_capacity(cap);
    cudaError_t e = ctx->scratch[0].reserve(npx * 4);                 // eig map
    if (e == cudaSuccess) e = ctx->scratch[1].reserve(256 + 3 * (size_t)BS_BINS * 4);   // max bits, count, n_out, largest bucket | histogram, starts, cursors
    if (e == cudaSuccess) e = ctx->scratch[2].reserve((size_t)cap2 * sizeof(Rec128));
    if (e == cudaSuccess) e = ctx->scratch[3].reserve(npx * 4);      // rank map
    if (e == cudaSuccess) e = ctx->scratch[4].reserve(2 * (size_t)cap + 64);   // status | neighbour counts
    if (e == cudaSuccess) e = ctx->pin[0].reserve(64);
    if (e != cudaSuccess) return ctx->fail(PMV_ERR_NOMEM, "gftt workspace", e);
    float *d_eig = ctx->scratch[0].as<float>();
    int *d_misc = ctx->scratch[1].as<int>();   // [0] max bits, [1] count, [2] n_out
    Rec128 *d_rec = ctx->scratch[2].as<Rec128>();
    int *d_rank = ctx->scratch[3].as<int>();
    unsigned char *d_status = ctx->scratch[4].as<unsigned char>();
    int *h_misc = ctx->pin[0].as<int>();

    PMV_CUDA_TRY(ctx, cudaMemsetAsync(d_misc, 0, 16, s));
    rc = run_mineig(ctx, v, d_eig, d_misc, s);
    if (rc) return rc;
    int n_cand = 0;
    {
        ProfScope ps(ctx, PMV_PHASE_SELECT, s);
        dim3 grid((v.rw + 31) / 32, (v.rh + GC_ROWS - 1) / GC_ROWS);
        int *d_hist = d_misc + 64, *d_start = d_hist + BS_BINS, *d_cursor = d_start + BS_BINS;
        const bool big = npx >= (size_t)1 << 19;    // small views: a few thousand candidates, the bitonic network is two launches
        if (big) PMV_CUDA_TRY(ctx, cudaMemsetAsync(d_hist, 0, sizeof(int) * 3 * (size_t)BS_BINS, s));
        gftt_candidates_kernel<<<grid, 256, 0, s>>>(d_eig, v.rh, v.rw, d_misc, quality, d_rec, d_misc + 1, cap, big ? d_hist : nullptr, d_misc + 3);
        PMV_LAUNCH_CHECK(ctx, "gftt_candidates_kernel");
        if (big) {
            if (ctx->attr_first(PMV_ATTR_BS_SCAN)) cudaFuncSetAttribute(bs_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BS_SCAN_SMEM);
            bs_scan_kernel<<<1, 1024, BS_SCAN_SMEM, s>>>(d_hist, d_start, d_misc + 3);
            PMV_LAUNCH_CHECK(ctx, "bs_scan_kernel");
        }
        PMV_CUDA_TRY(ctx, cudaMemcpyAsync(h_misc, d_misc, 16, cudaMemcpyDeviceToHost, s));
        PMV_CUDA_TRY(ctx, cudaStreamSynchronize(s));
        n_cand = h_misc[1] < cap ? h_misc[1] : cap;
        const int want = max_corners > 0 ? max_corners : n_cand;
        if (n_cand > 0) {
            if (big && n_cand >= 4 * SORT_CHUNK && h_misc[1] <= cap && h_misc[3] <= BS_MAXBUCKET) {
                // bucket sort: scatter by the top bits of the response, then rank inside the buckets (sort.cuh)
                e = ctx->scratch[6].reserve((size_t)n_cand * GF_NBMAX * 4 + 16);   // >= 16 B per record; reused for the neighbour lists below
                if (e != cudaSuccess) return ctx->fail(PMV_ERR_NOMEM, "gftt sort workspace", e);
                Rec128 *d_tmp = ctx->scratch[6].as<Rec128>();
                bs_scatter_kernel<<<(n_cand + 255) / 256, 256, 0, s>>>(d_rec, n_cand, d_misc, d_start, d_cursor, d_tmp);
                PMV_LAUNCH_CHECK(ctx, "bs_scatter_kernel");
                bs_rank_kernel<<<(n_cand + 255) / 256, 256, 0, s>>>(d_tmp, n_cand, d_misc, d_start, d_hist, d_rec);
                PMV_LAUNCH_CHECK(ctx, "bs_rank_kernel");
            } else {
                rc = sort_desc_128(ctx, d_rec, n_cand, s);
                if (rc) return rc;
            }
            PMV_CUDA_TRY(ctx, cudaMemsetAsync(d_rank, 0xff, npx * 4, s));
            PMV_CUDA_TRY(ctx, cudaMemsetAsync(d_status, 0, n_cand, s));
            rank_scatter_kernel<<<(n_cand + 255) / 256, 256, 0, s>>>(d_rec, n_cand, d_rank);
            PMV_LAUNCH_CHECK(ctx, "rank_scatter_kernel");
            e = ctx->scratch[5].reserve((size_t)want * 12 + 16);
            if (e != cudaSuccess) return ctx->fail(PMV_ERR_NOMEM, "gftt output", e);
            float *d_xy = ctx->scratch[5].as<float>();
            float *d_sc = d_xy + 2 * (size_t)want;
            e = ctx->scratch[6].reserve((size_t)n_cand * GF_NBMAX * 4 + 16);
            if (e != cudaSuccess) return ctx->fail(PMV_ERR_NOMEM, "gftt neighbour lists", e);
            int *d_nb = ctx->scratch[6].as<int>();
            unsigned char *d_nbc = d_status + cap;
            // Neighbour lists for the candidates the greedy pass is likely to reach (it stops after max_corners accepted
            // corners, normally within the first few multiples of that); the rest are marked "scan the window" (255) --
            // same result either way, and 80 % of the list work of a 4K frame is not done.
            const int n_lists = max_corners > 0 ? std::min(n_cand, 4 * max_corners) : n_cand;
            if (n_lists < n_cand) PMV_CUDA_TRY(ctx, cudaMemsetAsync(d_nbc + n_lists, 0xff, n_cand - n_lists, s));
            gftt_neighbors_kernel<<<(n_lists + 255) / 256, 256, 0, s>>>(d_rec, n_lists, v.rh, v.rw, d_rank, (float)min_dist, d_nb, d_nbc);
            PMV_LAUNCH_CHECK(ctx, "gftt_neighbors_kernel");
            gftt_select_kernel<<<1, 1024, 0, s>>>(d_rec, n_cand, v.rh, v.rw, d_rank, d_nb, d_nbc, d_status, (float)min_dist,
                                                  max_corners, d_xy, d_sc, d_misc + 2);
            PMV_LAUNCH_CHECK(ctx, "gftt_select_kernel");
            PMV_CUDA_TRY(ctx, cudaMemcpyAsync(h_misc, d_misc, 16, cudaMemcpyDeviceToHost, s));
            PMV_CUDA_TRY(ctx, cudaStreamSynchronize(s));
            *out_xy = d_xy; *out_sc = d_sc; *n_out = h_misc[2];
            return PMV_OK;
        }
    }
    *n_out = 0;
    return PMV_OK;
}

}  // namespace

int pmv_internal_gftt_device(pmv_ctx *ctx, const uint8_t *d_img, int pitch, int full_rows, int full_cols, int rx, int ry, int rw,
                             int rh, int max_corners, double quality, double min_dist, float **d_xy, float **d_score, int *n)
{
    ImgView v{d_img, pitch, 0, 0, full_rows, full_cols, rx, ry, rw, rh};
    v.word_ok = (((uintptr_t)d_img | (uintptr_t)pitch | (uintptr_t)rx) & 3) == 0;
    return gftt_run(ctx, v, max_corners, quality, min_dist, d_xy, d_score, n);
}

extern "C" {

PMV_API int pmv_min_eigen_val(pmv_ctx *ctx, const uint8_t *base, int full_rows, int full_cols, int step,
                              int roi_x, int roi_y, int roi_w, int roi_h, float *eig)
{
    if (!ctx) return PMV_ERR_INVALID;
    int rc = check_roi(ctx, base, full_rows, full_cols, step, roi_x, roi_y, roi_w, roi_h);
    if (rc) return rc;
    if (!eig) return ctx->fail(PMV_ERR_INVALID, "pmv_min_eigen_val: null output");
    cudaSetDevice(ctx->device);
    ImgView v;
    rc = upload_roi(ctx, base, full_rows, full_cols, step, roi_x, roi_y, roi_w, roi_h, &v);
    if (rc) return rc;
    size_t n = (size_t)roi_w * roi_h;
    cudaError_t e = ctx->scratch[0].reserve(n * 4);
    if (e == cudaSuccess) e = ctx->scratch[1].reserve(64);
    if (e != cudaSuccess) return ctx->fail(PMV_ERR_NOMEM, "eig map", e);
    rc = run_mineig(ctx, v, ctx->scratch[0].as<float>(), ctx->scratch[1].as<int>(), ctx->stream);
    if (rc) return rc;
    PMV_CUDA_TRY(ctx, cudaMemcpyAsync(eig, ctx->scratch[0].p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PMV_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return PMV_OK;
}

PMV_API int pmv_min_eigen_val_batched_dev(pmv_ctx *ctx, const uint8_t *d_imgs, int batch, size_t img_stride, int rows, int cols,
                                          int step, float *d_eig, float *d_max)
{
    if (!ctx) return PMV_ERR_INVALID;
    if (!d_imgs || !d_eig || !d_max || batch <= 0 || rows <= 0 || cols <= 0 || step < cols || img_stride < (size_t)rows * step)
        return ctx->fail(PMV_ERR_INVALID, "pmv_min_eigen_val_batched_dev: bad argument");
    cudaSetDevice(ctx->device);
    ImgView v{d_imgs, step, 0, 0, rows, cols, 0, 0, cols, rows};
    v.word_ok = (((uintptr_t)d_imgs | (uintptr_t)step | (uintptr_t)img_stride) & 3) == 0;
    // the running maxima are the bits of non-negative floats: d_max doubles as the int buffer of the kernels
    return run_mineig(ctx, v, d_eig, reinterpret_cast<int *>(d_max), ctx->stream, batch, img_stride, (size_t)rows * cols);
}

PMV_API int pmv_shitomasi_response_batched_dev(pmv_ctx *ctx, const uint8_t *d_imgs, int batch, size_t img_stride, int rows,
                                               int cols, int step, int signed_quirk, double *d_R, double *d_max)
{
    if (!ctx) return PMV_ERR_INVALID;
    if (!d_imgs || !d_R || !d_max || batch <= 0 || rows <= 0 || cols <= 0 || step < cols || img_stride < (size_t)rows * step)
        return ctx->fail(PMV_ERR_INVALID, "pmv_shitomasi_response_batched_dev: bad argument");
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    ImgView v{d_imgs, step, 0, 0, rows, cols, 0, 0, cols, rows};
    v.word_ok = (((uintptr_t)d_imgs | (uintptr_t)step | (uintptr_t)img_stride) & 3) == 0;
    ProfScope ps(ctx, PMV_PHASE_RESPONSE, s);
    PMV_CUDA_TRY(ctx, cudaMemsetAsync(d_max, 0, sizeof(double) * batch, s));
    return run_shitomasi(ctx, v, signed_quirk, d_R, reinterpret_cast<unsigned long long *>(d_max), s, batch, img_stride, (size_t)rows * cols);
}

PMV_API int pmv_gftt(pmv_ctx *ctx, const uint8_t *base, int full_rows, int full_cols, int step,
                     int roi_x, int roi_y, int roi_w, int roi_h, int max_corners, double quality,
                     double min_dist, int block_size, int ksize, float *xy, float *score, int *n_out)
{
    if (!ctx) return PMV_ERR_INVALID;
    int rc = check_roi(ctx, base, full_rows, full_cols, step, roi_x, roi_y, roi_w, roi_h);
    if (rc) return rc;
    if (!n_out || quality <= 0 || min_dist < 0) return ctx->fail(PMV_ERR_INVALID, "pmv_gftt: bad argument");
    if (block_size != 3 || ksize != 3)
        return ctx->fail(PMV_ERR_UNSUPPORTED, "pmv_gftt: only blockSize 3 / ksize 3 (the reference's call)");
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    ImgView v;
    rc = upload_roi(ctx, base, full_rows, full_cols, step, roi_x, roi_y, roi_w, roi_h, &v);
    if (rc) return rc;
    float *d_xy = nullptr, *d_sc = nullptr;
    int n = 0;
    rc = gftt_run(ctx, v, max_corners, quality, min_dist, &d_xy, &d_sc, &n);
    if (rc) return rc;
    if (n > 0) {
        if (xy) PMV_CUDA_TRY(ctx, cudaMemcpyAsync(xy, d_xy, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
        if (score) PMV_CUDA_TRY(ctx, cudaMemcpyAsync(score, d_sc, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
        PMV_CUDA_TRY(ctx, cudaStreamSynchronize(s));
    }
    *n_out = n;
    return PMV_OK;
}

PMV_API int pmv_gftt_dev(pmv_ctx *ctx, const uint8_t *d_base, int full_rows, int full_cols, int step,
                         int roi_x, int roi_y, int roi_w, int roi_h, int max_corners, double quality,
                         double min_dist, float *d_xy_out, float *d_score_out, int *n_out)
{
    if (!ctx) return PMV_ERR_INVALID;
    int rc = check_roi(ctx, d_base, full_rows, full_cols, step, roi_x, roi_y, roi_w, roi_h);
    if (rc) return rc;
    if (!n_out || quality <= 0 || min_dist < 0 || max_corners <= 0) return ctx->fail(PMV_ERR_INVALID, "pmv_gftt_dev: bad argument");
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    // the resident image is read in place: the view starts at parent pixel (0, 0)
    ImgView v{d_base, step, 0, 0, full_rows, full_cols, roi_x, roi_y, roi_w, roi_h};
    v.word_ok = (((uintptr_t)d_base | (uintptr_t)step | (uintptr_t)roi_x) & 3) == 0;
    float *d_xy = nullptr, *d_sc = nullptr;
    int n = 0;
    rc = gftt_run(ctx, v, max_corners, quality, min_dist, &d_xy, &d_sc, &n);
    if (rc) return rc;
    if (n > 0) {
        if (d_xy_out) PMV_CUDA_TRY(ctx, cudaMemcpyAsync(d_xy_out, d_xy, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
        if (d_score_out) PMV_CUDA_TRY(ctx, cudaMemcpyAsync(d_score_out, d_sc, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
    }
    *n_out = n;
    return PMV_OK;
}

PMV_API int pmv_shitomasi_response(pmv_ctx *ctx, const uint8_t *img, int rows, int cols, int step,
                                   int signed_quirk, double *R)
{
    if (!ctx) return PMV_ERR_INVALID;
    int rc = check_roi(ctx, img, rows, cols, step, 0, 0, cols, rows);
    if (rc) return rc;
    if (!R) return ctx->fail(PMV_ERR_INVALID, "pmv_shitomasi_response: null output");
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    ImgView v;
    rc = upload_roi(ctx, img, rows, cols, step, 0, 0, cols, rows, &v);
    if (rc) return rc;
    size_t n = (size_t)rows * cols;
    cudaError_t e = ctx->scratch[0].reserve(n * 8);
    if (e == cudaSuccess) e = ctx->scratch[1].reserve(256);
    if (e != cudaSuccess) return ctx->fail(PMV_ERR_NOMEM, "shitomasi map", e);
    PMV_CUDA_TRY(ctx, cudaMemsetAsync(ctx->scratch[1].p, 0, 16, s));
    rc = run_shitomasi(ctx, v, signed_quirk, ctx->scratch[0].as<double>(), ctx->scratch[1].as<unsigned long long>(), s);
    if (rc) return rc;
    PMV_CUDA_TRY(ctx, cudaMemcpyAsync(R, ctx->scratch[0].p, n * 8, cudaMemcpyDeviceToHost, s));
    PMV_CUDA_TRY(ctx, cudaStreamSynchronize(s));
    return PMV_OK;
}

PMV_API int pmv_shitomasi(pmv_ctx *ctx, const uint8_t *img, int rows, int cols, int step, int max_feats,
                          double quality, int signed_quirk, int *col, int *row, double *score, int *n_out)
{
    if (!ctx) return PMV_ERR_INVALID;
    int rc = check_roi(ctx, img, rows, cols, step, 0, 0, cols, rows);
    if (rc) return rc;
    if (!n_out || max_feats < 0) return ctx->fail(PMV_ERR_INVALID, "pmv_shitomasi: bad argument");
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    ImgView v;
    rc = upload_roi(ctx, img, rows, cols, step, 0, 0, cols, rows, &v);
    if (rc) return rc;
    const size_t npx = (size_t)rows * cols;
    const int cap2 = sort_capacity((int)npx);
    cudaError_t e = ctx->scratch[0].reserve(npx * 8);
    if (e == cudaSuccess) e = ctx->scratch[1].reserve(256);
    if (e == cudaSuccess) e = ctx->scratch[2].reserve((size_t)cap2 * sizeof(Rec128));
    if (e == cudaSuccess) e = ctx->pin[0].reserve(64);
    if (e != cudaSuccess) return ctx->fail(PMV_ERR_NOMEM, "shitomasi workspace", e);
    double *d_R = ctx->scratch[0].as<double>();
    unsigned long long *d_max = ctx->scratch[1].as<unsigned long long>();
    int *d_count = reinterpret_cast<int *>(d_max + 1);
    Rec128 *d_rec = ctx->scratch[2].as<Rec128>();
    int *h_misc = ctx->pin[0].as<int>();
    PMV_CUDA_TRY(ctx, cudaMemsetAsync(d_max, 0, 16, s));
    {
        ProfScope ps(ctx, PMV_PHASE_RESPONSE, s);
        rc = run_shitomasi(ctx, v, signed_quirk, d_R, d_max, s);
        if (rc) return rc;
    }
    ProfScope ps(ctx, PMV_PHASE_SELECT, s);
    shitomasi_candidates_kernel<<<(int)((npx + 255) / 256), 256, 0, s>>>(d_R, (int)npx, d_max, quality, d_rec, d_count, (int)npx);
    PMV_LAUNCH_CHECK(ctx, "shitomasi_candidates_kernel");
    PMV_CUDA_TRY(ctx, cudaMemcpyAsync(h_misc, d_count, 4, cudaMemcpyDeviceToHost, s));
    PMV_CUDA_TRY(ctx, cudaStreamSynchronize(s));
    int n_cand = h_misc[0];
    int n = n_cand < max_feats ? n_cand : max_feats;
    if (n > 0) {
        rc = sort_desc_128(ctx, d_rec, n_cand, s);
        if (rc) return rc;
        e = ctx->scratch[5].reserve((size_t)n * 16 + 16);
        if (e != cudaSuccess) return ctx->fail(PMV_ERR_NOMEM, "shitomasi output", e);
        double *d_sc = ctx->scratch[5].as<double>();
        int *d_col = reinterpret_cast<int *>(d_sc + n), *d_row = d_col + n;
        shitomasi_emit_kernel<<<(n + 255) / 256, 256, 0, s>>>(d_rec, n, cols, d_col, d_row, d_sc);
        PMV_LAUNCH_CHECK(ctx, "shitomasi_emit_kernel");
        if (col) PMV_CUDA_TRY(ctx, cudaMemcpyAsync(col, d_col, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
        if (row) PMV_CUDA_TRY(ctx, cudaMemcpyAsync(row, d_row, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
        if (score) PMV_CUDA_TRY(ctx, cudaMemcpyAsync(score, d_sc, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
        PMV_CUDA_TRY(ctx, cudaStreamSynchronize(s));
    }
    *n_out = n;
    return PMV_OK;
}

}  // extern "C"
