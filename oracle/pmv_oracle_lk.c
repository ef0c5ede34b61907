/*
 * oracle/pmv_oracle_lk.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Plain-C restatement of the image-pyramid build, Scharr derivative and the
 * per-feature pyramidal Lucas-Kanade solve that the reference reaches through
 *   /root/reference/OpenCVLucasKanadeFM.cpp:15
 *     cv::calcOpticalFlowPyrLK(src.bw, next.bw, prev, next, status, err,
 *                              cv::Size(win_size, win_size), pyr_size);
 * The arithmetic lives in OpenCV (un-vendored third party; the reference pins
 * "3.3 or later", README.md:7).  This file restates OpenCV's published
 * algorithm and is PINNED against the cv2 4.13.0 build
 * of the very same kernels: tests/test_oracle_lk.py compares every function
 * here with cv2.pyrDown / cv2.buildOpticalFlowPyramid / cv2.calcOpticalFlowPyrLK
 * and with the committed fixtures under tests/golden/.
 *
 * Nothing under practical-multi-view_b200/ may link or call this file; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg do.
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

static inline int reflect101(int p, int len)
{
    /* BORDER_REFLECT_101: gfedcb|abcdefgh|gfedcba */
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    }
    return p;
}

/* ---- A.1  pyrDown: separable [1 4 6 4 1], reflect-101, (s + 128) >> 8 ------ */
ORC_API void orc_pyr_down(const uint8_t *src, int rows, int cols, int step,
                          uint8_t *dst, int dstep)
{
    int drows = (rows + 1) / 2, dcols = (cols + 1) / 2;
    for (int y = 0; y < drows; y++) {
        for (int x = 0; x < dcols; x++) {
            int acc = 0;
            static const int k[5] = {1, 4, 6, 4, 1};
            for (int dy = -2; dy <= 2; dy++) {
                int sy = reflect101(2 * y + dy, rows);
                int racc = 0;
                for (int dx = -2; dx <= 2; dx++) {
                    int sx = reflect101(2 * x + dx, cols);
                    racc += k[dx + 2] * src[(size_t)sy * step + sx];
                }
                acc += k[dy + 2] * racc;
            }
            dst[(size_t)y * dstep + x] = (uint8_t)((acc + 128) >> 8);
        }
    }
}

/* Effective level count of cv::buildOpticalFlowPyramid (SURVEY Appx A.1):
 * after producing level l, stop if the next size is <= the window. */
ORC_API int orc_pyr_levels(int rows, int cols, int win_w, int win_h, int max_level)
{
    int w = cols, h = rows, level = 0;
    for (level = 0; level <= max_level; level++) {
        int nw = (w + 1) / 2, nh = (h + 1) / 2;
        if (level == max_level) break;
        if (nw <= win_w || nh <= win_h) break;
        w = nw; h = nh;
    }
    return level;
}

/* ---- A.2  calcSharrDeriv: int16 x2 interleaved [Ix, Iy], reflect-101 ------- */
ORC_API void orc_scharr(const uint8_t *src, int rows, int cols, int step,
                        int16_t *dst /* rows*cols*2, packed */)
{
    for (int y = 0; y < rows; y++) {
        int y0 = reflect101(y - 1, rows), y2 = reflect101(y + 1, rows);
        for (int x = 0; x < cols; x++) {
            int xm = reflect101(x - 1, cols), xp = reflect101(x + 1, cols);
#define PX(yy, xx) ((int)src[(size_t)(yy) * step + (xx)])
            int t0m = 3 * (PX(y0, xm) + PX(y2, xm)) + 10 * PX(y, xm);
            int t0p = 3 * (PX(y0, xp) + PX(y2, xp)) + 10 * PX(y, xp);
            int t1m = PX(y2, xm) - PX(y0, xm);
            int t1c = PX(y2, x) - PX(y0, x);
            int t1p = PX(y2, xp) - PX(y0, xp);
#undef PX
            dst[((size_t)y * cols + x) * 2 + 0] = (int16_t)(t0p - t0m);
            dst[((size_t)y * cols + x) * 2 + 1] = (int16_t)(3 * (t1m + t1p) + 10 * t1c);
        }
    }
}

/* ---- A.3  LKTrackerInvoker -------------------------------------------------- */
typedef struct {
    const uint8_t *img; /* level image, no border; reads outside go through reflect-101 */
    int rows, cols, step;
} orc_level;

static inline int lvl_px(const orc_level *L, int y, int x)
{
    return L->img[(size_t)reflect101(y, L->rows) * L->step + reflect101(x, L->cols)];
}

/* derivative image has a ZERO border (BORDER_CONSTANT) outside the image */
static inline int drv_px(const int16_t *d, int rows, int cols, int y, int x, int c)
{
    if (y < 0 || y >= rows || x < 0 || x >= cols) return 0;
    return d[((size_t)y * cols + x) * 2 + c];
}

static inline int round_half_even(float v)
{
    return (int)lrintf(v); /* default FE_TONEAREST == cvRound */
}

#define DESCALE(x, n) (((x) + (1 << ((n)-1))) >> (n))

static void lk_level(const orc_level *I, const orc_level *J, const int16_t *dI,
                     const float *prev_xy, float *next_xy, uint8_t *status, float *err,
                     int n, int win_w, int win_h, int level, int max_level,
                     int max_count, double eps2, int flags, float min_eig_thr,
                     int16_t *Iwin, int16_t *dwin)
{
    const float halfx = (win_w - 1) * 0.5f, halfy = (win_h - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);
    for (int i = 0; i < n; i++) {
        float sc = (float)(1. / (1 << level));
        float px = prev_xy[2 * i] * sc, py = prev_xy[2 * i + 1] * sc;
        float nx, ny;
        if (level == max_level) {
            if (flags & 4) { nx = next_xy[2 * i] * sc; ny = next_xy[2 * i + 1] * sc; }
            else { nx = px; ny = py; }
        } else {
            nx = next_xy[2 * i] * 2.f; ny = next_xy[2 * i + 1] * 2.f;
        }
        next_xy[2 * i] = nx; next_xy[2 * i + 1] = ny;

        px -= halfx; py -= halfy;
        int ipx = (int)floorf(px), ipy = (int)floorf(py);
        if (ipx < -win_w || ipx >= I->cols || ipy < -win_h || ipy >= I->rows) {
            if (level == 0) { status[i] = 0; err[i] = 0; }
            continue;
        }
        float a = px - ipx, b = py - ipy;
        int iw00 = round_half_even((1.f - a) * (1.f - b) * (1 << 14));
        int iw01 = round_half_even(a * (1.f - b) * (1 << 14));
        int iw10 = round_half_even((1.f - a) * b * (1 << 14));
        int iw11 = (1 << 14) - iw00 - iw01 - iw10;
        float A11 = 0, A12 = 0, A22 = 0;
        for (int y = 0; y < win_h; y++) {
            for (int x = 0; x < win_w; x++) {
                int yy = y + ipy, xx = x + ipx;
                int ival = DESCALE(lvl_px(I, yy, xx) * iw00 + lvl_px(I, yy, xx + 1) * iw01 +
                                   lvl_px(I, yy + 1, xx) * iw10 + lvl_px(I, yy + 1, xx + 1) * iw11, 9);
                int ixval = DESCALE(drv_px(dI, I->rows, I->cols, yy, xx, 0) * iw00 +
                                    drv_px(dI, I->rows, I->cols, yy, xx + 1, 0) * iw01 +
                                    drv_px(dI, I->rows, I->cols, yy + 1, xx, 0) * iw10 +
                                    drv_px(dI, I->rows, I->cols, yy + 1, xx + 1, 0) * iw11, 14);
                int iyval = DESCALE(drv_px(dI, I->rows, I->cols, yy, xx, 1) * iw00 +
                                    drv_px(dI, I->rows, I->cols, yy, xx + 1, 1) * iw01 +
                                    drv_px(dI, I->rows, I->cols, yy + 1, xx, 1) * iw10 +
                                    drv_px(dI, I->rows, I->cols, yy + 1, xx + 1, 1) * iw11, 14);
                Iwin[y * win_w + x] = (int16_t)ival;
                dwin[(y * win_w + x) * 2] = (int16_t)ixval;
                dwin[(y * win_w + x) * 2 + 1] = (int16_t)iyval;
                A11 += (float)(ixval * ixval);
                A12 += (float)(ixval * iyval);
                A22 += (float)(iyval * iyval);
            }
        }
        A11 *= FLT_SCALE; A12 *= FLT_SCALE; A22 *= FLT_SCALE;
        float D = A11 * A22 - A12 * A12;
        float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) /
                       (2 * win_w * win_h);
        if ((flags & 8) != 0) err[i] = minEig;
        if (minEig < min_eig_thr || D < FLT_EPSILON) {
            if (level == 0) status[i] = 0;
            continue;
        }
        D = 1.f / D;
        nx -= halfx; ny -= halfy;
        float pdx = 0, pdy = 0;
        for (int j = 0; j < max_count; j++) {
            int inx = (int)floorf(nx), iny = (int)floorf(ny);
            if (inx < -win_w || inx >= J->cols || iny < -win_h || iny >= J->rows) {
                if (level == 0) status[i] = 0;
                break;
            }
            a = nx - inx; b = ny - iny;
            iw00 = round_half_even((1.f - a) * (1.f - b) * (1 << 14));
            iw01 = round_half_even(a * (1.f - b) * (1 << 14));
            iw10 = round_half_even((1.f - a) * b * (1 << 14));
            iw11 = (1 << 14) - iw00 - iw01 - iw10;
            float b1 = 0, b2 = 0;
            for (int y = 0; y < win_h; y++) {
                for (int x = 0; x < win_w; x++) {
                    int yy = y + iny, xx = x + inx;
                    int diff = DESCALE(lvl_px(J, yy, xx) * iw00 + lvl_px(J, yy, xx + 1) * iw01 +
                                       lvl_px(J, yy + 1, xx) * iw10 + lvl_px(J, yy + 1, xx + 1) * iw11, 9) -
                               Iwin[y * win_w + x];
                    b1 += (float)(diff * dwin[(y * win_w + x) * 2]);
                    b2 += (float)(diff * dwin[(y * win_w + x) * 2 + 1]);
                }
            }
            b1 *= FLT_SCALE; b2 *= FLT_SCALE;
            float dx = (float)((A12 * b2 - A22 * b1) * D);
            float dy = (float)((A12 * b1 - A11 * b2) * D);
            nx += dx; ny += dy;
            next_xy[2 * i] = nx + halfx; next_xy[2 * i + 1] = ny + halfy;
            if ((double)dx * dx + (double)dy * dy <= eps2) break;
            if (j > 0 && fabsf(dx + pdx) < 0.01 && fabsf(dy + pdy) < 0.01) {
                next_xy[2 * i] -= dx * 0.5f; next_xy[2 * i + 1] -= dy * 0.5f;
                break;
            }
            pdx = dx; pdy = dy;
        }
        if (status[i] && level == 0 && (flags & 8) == 0) {
            float fx = next_xy[2 * i] - halfx, fy = next_xy[2 * i + 1] - halfy;
            int inx = (int)floorf(fx), iny = (int)floorf(fy);
            if (inx < -win_w || inx >= J->cols || iny < -win_h || iny >= J->rows) {
                status[i] = 0;
                continue;
            }
            float aa = fx - inx, bb = fy - iny;
            iw00 = round_half_even((1.f - aa) * (1.f - bb) * (1 << 14));
            iw01 = round_half_even(aa * (1.f - bb) * (1 << 14));
            iw10 = round_half_even((1.f - aa) * bb * (1 << 14));
            iw11 = (1 << 14) - iw00 - iw01 - iw10;
            float errval = 0.f;
            for (int y = 0; y < win_h; y++)
                for (int x = 0; x < win_w; x++) {
                    int yy = y + iny, xx = x + inx;
                    int diff = DESCALE(lvl_px(J, yy, xx) * iw00 + lvl_px(J, yy, xx + 1) * iw01 +
                                       lvl_px(J, yy + 1, xx) * iw10 + lvl_px(J, yy + 1, xx + 1) * iw11, 9) -
                               Iwin[y * win_w + x];
                    errval += fabsf((float)diff);
                }
            err[i] = errval * 1.f / (32 * win_w * win_h);
        }
    }
}

/* == cv::calcOpticalFlowPyrLK(prev, next, prevPts, nextPts, status, err,
 *                             Size(win_w,win_h), max_level, {COUNT+EPS,max_count,eps}, flags, min_eig)
 * flags: 4 = OPTFLOW_USE_INITIAL_FLOW, 8 = OPTFLOW_LK_GET_MIN_EIGENVALS.
 * returns the effective max level. */
ORC_API int orc_lk_track(const uint8_t *prev, const uint8_t *next, int rows, int cols, int step,
                         const float *prev_xy, int n, int win_w, int win_h, int max_level,
                         int max_count, double eps, int flags, double min_eig_thr,
                         float *next_xy, uint8_t *status, float *err)
{
    if (max_count < 0) max_count = 0;
    if (max_count > 100) max_count = 100;
    if (eps < 0) eps = 0;
    if (eps > 10) eps = 10;
    double eps2 = eps * eps;
    int L = orc_pyr_levels(rows, cols, win_w, win_h, max_level);
    orc_level P[16], N[16];
    uint8_t *bufP[16] = {0}, *bufN[16] = {0};
    P[0].img = prev; P[0].rows = rows; P[0].cols = cols; P[0].step = step;
    N[0].img = next; N[0].rows = rows; N[0].cols = cols; N[0].step = step;
    for (int l = 1; l <= L; l++) {
        int r = (P[l - 1].rows + 1) / 2, c = (P[l - 1].cols + 1) / 2;
        bufP[l] = (uint8_t *)malloc((size_t)r * c);
        bufN[l] = (uint8_t *)malloc((size_t)r * c);
        orc_pyr_down(P[l - 1].img, P[l - 1].rows, P[l - 1].cols, P[l - 1].step, bufP[l], c);
        orc_pyr_down(N[l - 1].img, N[l - 1].rows, N[l - 1].cols, N[l - 1].step, bufN[l], c);
        P[l].img = bufP[l]; P[l].rows = r; P[l].cols = c; P[l].step = c;
        N[l].img = bufN[l]; N[l].rows = r; N[l].cols = c; N[l].step = c;
    }
    for (int i = 0; i < n; i++) { status[i] = 1; err[i] = 0; }
    int16_t *Iwin = (int16_t *)malloc(sizeof(int16_t) * win_w * win_h);
    int16_t *dwin = (int16_t *)malloc(sizeof(int16_t) * win_w * win_h * 2);
    int16_t *dI = (int16_t *)malloc(sizeof(int16_t) * 2 * (size_t)rows * cols);
    for (int l = L; l >= 0; l--) {
        orc_scharr(P[l].img, P[l].rows, P[l].cols, P[l].step, dI);
        lk_level(&P[l], &N[l], dI, prev_xy, next_xy, status, err, n, win_w, win_h, l, L,
                 max_count, eps2, flags, (float)min_eig_thr, Iwin, dwin);
    }
    free(Iwin); free(dwin); free(dI);
    for (int l = 1; l <= L; l++) { free(bufP[l]); free(bufN[l]); }
    return L;
}
