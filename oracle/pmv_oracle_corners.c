/*
 * oracle/pmv_oracle_corners.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Restates the corner detectors on the reference's hot path:
 *   orc_min_eigen_val / orc_gftt : cv::goodFeaturesToTrack(bw, corners, max, 0.01, 5, Mat(), 3, 3, false, .04)
 *                                  called at /root/reference/OpenCVGoodFeatureExtractor.cpp:7
 *                                  (third-party OpenCV; published algorithm, SURVEY Appx B.1/B.2)
 *   orc_shitomasi               : /root/reference/ShiTomasiFeatureExtractor.cpp:5-75 on top of
 *                                  /root/reference/Frame.cpp:58-86 (computeSpatialGradient) and
 *                                  Frame.cpp:119-138 (computeHarrisMatrix) -- the reference's own code
 *   orc_fast                    : cv::FAST(bw, kp, 10, true) at /root/reference/OpenCVFASTFeatureExtractor.cpp:8
 * Pinned by tests/test_oracle_corners.py against cv2 4.13.0 (cornerMinEigenVal, goodFeaturesToTrack,
 * FastFeatureDetector, blur) and the fixtures in tests/golden/.
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

static inline int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

/* ---- B.1 cornerMinEigenVal(src, 3, 3) on an ROI of a parent image -------------------------
 * C++ semantics of a cv::Mat sub-view (Frame::regionOfInterest, Frame.cpp:98-99): Sobel reads
 * parent pixels beyond the ROI edge and reflects at the PARENT's edges; the covariance image is a
 * fresh Mat, so the 3x3 box sum reflects at the ROI's own edges. */
ORC_API void orc_min_eigen_val(const uint8_t *img, int full_rows, int full_cols, int step,
                               int rx, int ry, int rw, int rh, float *eig /* rh*rw */)
{
    const double scale = 1.0 / (4.0 * 3.0 * 255.0); /* 1/(2^(ksize-1) * blockSize * 255) */
    const float k0 = (float)scale, k1 = (float)(2 * scale);
    float *dx = (float *)malloc(sizeof(float) * rw * rh), *dy = (float *)malloc(sizeof(float) * rw * rh);
#define PX(y, x) ((float)img[(size_t)reflect101((y), full_rows) * step + reflect101((x), full_cols)])
    for (int y = 0; y < rh; y++)
        for (int x = 0; x < rw; x++) {
            int gy = ry + y, gx = rx + x;
            /* Dx: row filter [-1 0 1] (exact), column filter [k0 k1 k0] */
            float ra = PX(gy - 1, gx + 1) - PX(gy - 1, gx - 1);
            float rc = PX(gy, gx + 1) - PX(gy, gx - 1);
            float rb = PX(gy + 1, gx + 1) - PX(gy + 1, gx - 1);
            dx[y * rw + x] = fmaf(k0, ra + rb, k1 * rc);
            /* Dy: row filter [k0 k1 k0], column filter [-1 0 1] */
            float top = fmaf(k0, PX(gy - 1, gx - 1) + PX(gy - 1, gx + 1), k1 * PX(gy - 1, gx));
            float bot = fmaf(k0, PX(gy + 1, gx - 1) + PX(gy + 1, gx + 1), k1 * PX(gy + 1, gx));
            dy[y * rw + x] = bot - top;
        }
#undef PX
    for (int y = 0; y < rh; y++)
        for (int x = 0; x < rw; x++) {
            double s0 = 0, s1 = 0, s2 = 0; /* boxFilter(normalize=false): float in, double sum */
            for (int j = -1; j <= 1; j++)
                for (int i = -1; i <= 1; i++) {
                    int yy = reflect101(y + j, rh), xx = reflect101(x + i, rw);
                    float a = dx[yy * rw + xx], b = dy[yy * rw + xx];
                    s0 += (double)(a * a); s1 += (double)(a * b); s2 += (double)(b * b);
                }
            float a = (float)s0 * 0.5f, b = (float)s1, c = (float)s2 * 0.5f;
            eig[y * rw + x] = (a + c) - sqrtf((a - c) * (a - c) + b * b);
        }
    free(dx); free(dy);
}

typedef struct { float v; int idx; } cand_t;
static int cand_cmp(const void *pa, const void *pb)
{
    const cand_t *a = (const cand_t *)pa, *b = (const cand_t *)pb;
    if (a->v > b->v) return -1;
    if (a->v < b->v) return 1;
    return (a->idx > b->idx) ? -1 : (a->idx < b->idx);  /* ties: higher address first */
}

/* ---- B.2 goodFeaturesToTrack selection on a response map ------------------------------ */
ORC_API int orc_gftt_select(const float *eig_in, int rows, int cols, int max_corners, double quality,
                            double min_dist, float *xy /* cap*2 */, float *score, int cap)
{
    size_t n = (size_t)rows * cols;
    float *eig = (float *)malloc(sizeof(float) * n);
    memcpy(eig, eig_in, sizeof(float) * n);
    double maxv = -DBL_MAX;
    for (size_t i = 0; i < n; i++) if (eig[i] > maxv) maxv = eig[i];
    float thr = (float)(maxv * quality);
    for (size_t i = 0; i < n; i++) if (!(eig[i] > thr)) eig[i] = 0.f; /* THRESH_TOZERO */
    cand_t *c = (cand_t *)malloc(sizeof(cand_t) * (n ? n : 1));
    int nc = 0;
    for (int y = 1; y < rows - 1; y++)
        for (int x = 1; x < cols - 1; x++) {
            float v = eig[y * cols + x];
            if (v == 0.f) continue;
            float m = v; /* dilate 3x3 */
            for (int j = -1; j <= 1; j++)
                for (int i = -1; i <= 1; i++) {
                    float t = eig[(y + j) * cols + x + i];
                    if (t > m) m = t;
                }
            if (v == m) { c[nc].v = v; c[nc].idx = y * cols + x; nc++; }
        }
    qsort(c, nc, sizeof(cand_t), cand_cmp);
    int out = 0;
    if (min_dist >= 1) {
        int cell = (int)lrint(min_dist); /* cvRound */
        int gw = (cols + cell - 1) / cell, gh = (rows + cell - 1) / cell;
        /* grid of linked lists */
        int *head = (int *)malloc(sizeof(int) * gw * gh), *nxt = (int *)malloc(sizeof(int) * (nc ? nc : 1));
        int *ax = (int *)malloc(sizeof(int) * (nc ? nc : 1)), *ay = (int *)malloc(sizeof(int) * (nc ? nc : 1));
        for (int i = 0; i < gw * gh; i++) head[i] = -1;
        double md2 = min_dist * min_dist;
        for (int i = 0; i < nc; i++) {
            int y = c[i].idx / cols, x = c[i].idx - y * cols;
            int xc = x / cell, yc = y / cell;
            int x1 = xc - 1 < 0 ? 0 : xc - 1, y1 = yc - 1 < 0 ? 0 : yc - 1;
            int x2 = xc + 1 > gw - 1 ? gw - 1 : xc + 1, y2 = yc + 1 > gh - 1 ? gh - 1 : yc + 1;
            int good = 1;
            for (int yy = y1; yy <= y2 && good; yy++)
                for (int xx = x1; xx <= x2 && good; xx++)
                    for (int k = head[yy * gw + xx]; k >= 0; k = nxt[k]) {
                        double ddx = x - ax[k], ddy = y - ay[k];
                        if (ddx * ddx + ddy * ddy < md2) { good = 0; break; }
                    }
            if (good) {
                ax[out] = x; ay[out] = y; nxt[out] = head[yc * gw + xc]; head[yc * gw + xc] = out;
                if (out < cap) { xy[2 * out] = (float)x; xy[2 * out + 1] = (float)y; score[out] = c[i].v; }
                out++;
                if (max_corners > 0 && out == max_corners) break;
            }
        }
        free(head); free(nxt); free(ax); free(ay);
    } else {
        for (int i = 0; i < nc; i++) {
            int y = c[i].idx / cols, x = c[i].idx - y * cols;
            if (out < cap) { xy[2 * out] = (float)x; xy[2 * out + 1] = (float)y; score[out] = c[i].v; }
            out++;
            if (max_corners > 0 && out == max_corners) break;
        }
    }
    free(c); free(eig);
    return out;
}

ORC_API int orc_gftt(const uint8_t *img, int full_rows, int full_cols, int step, int rx, int ry, int rw, int rh,
                     int max_corners, double quality, double min_dist, float *xy, float *score, int cap)
{
    float *eig = (float *)malloc(sizeof(float) * rw * rh);
    orc_min_eigen_val(img, full_rows, full_cols, step, rx, ry, rw, rh, eig);
    int n = orc_gftt_select(eig, rh, rw, max_corners, quality, min_dist, xy, score, cap);
    free(eig);
    return n;
}

/* ---- B.3 the reference's own ShiTomasiFeatureExtractor ---------------------------------- */
/* response map (rows*cols doubles). signed_quirk=1 reproduces Frame.cpp:65-67 (u8 read as schar). */
ORC_API void orc_shitomasi_response(const uint8_t *img, int rows, int cols, int step, int signed_quirk, double *R)
{
    size_t n = (size_t)rows * cols;
    double *gx = (double *)calloc(n, sizeof(double)), *gy = (double *)calloc(n, sizeof(double));
    double *H = (double *)malloc(sizeof(double) * n * 3);
#define S(y, x) (signed_quirk ? (double)(int8_t)img[(size_t)(y) * step + (x)] : (double)img[(size_t)(y) * step + (x)])
    for (int r = 1; r < rows - 1; r++)            /* Frame.cpp:63-84, interior only */
        for (int c = 1; c < cols - 1; c++) {
            gx[r * cols + c] = 1. / 2. * S(r, c + 1) - 1. / 2. * S(r, c - 1);
            gy[r * cols + c] = 1. / 2. * S(r + 1, c) - 1. / 2. * S(r - 1, c);
        }
#undef S
    for (size_t i = 0; i < n; i++) {              /* Frame.cpp:125-134: Ixx, Iyy, Ixy channels */
        H[3 * i] = gx[i] * gx[i]; H[3 * i + 1] = gy[i] * gy[i]; H[3 * i + 2] = gx[i] * gy[i];
    }
    memset(R, 0, sizeof(double) * n);
    for (int r = 0; r < rows; r++)
        for (int c = 0; c < cols - 1; c++) {      /* ShiTomasiFeatureExtractor.cpp:58: last column stays 0 */
            double s[3] = {0, 0, 0};              /* cv::blur 3x3 normalised, BORDER_REFLECT_101 (Frame.cpp:136) */
            for (int j = -1; j <= 1; j++)
                for (int i = -1; i <= 1; i++) {
                    size_t q = (size_t)reflect101(r + j, rows) * cols + reflect101(c + i, cols);
                    s[0] += H[3 * q]; s[1] += H[3 * q + 1]; s[2] += H[3 * q + 2];
                }
            double Ixx = s[0] * (1.0 / 9), Iyy = s[1] * (1.0 / 9), Ixy = s[2] * (1.0 / 9);
            double B = -Ixx - Iyy;
            double C = Ixx * Iyy - pow(Ixy, 2);
            double l1 = (-B + sqrt(pow(B, 2) - 4 * C)) / 2;
            double l2 = (-B - sqrt(pow(B, 2) - 4 * C)) / 2;
            R[r * cols + c] = l1 < l2 ? l1 : l2;  /* std::min: NaN propagates like the reference */
        }
    free(gx); free(gy); free(H);
}

typedef struct { double v; int idx; } dcand_t;
static int dcand_cmp(const void *pa, const void *pb)
{
    const dcand_t *a = (const dcand_t *)pa, *b = (const dcand_t *)pb;
    if (a->v > b->v) return -1;
    if (a->v < b->v) return 1;
    return (a->idx < b->idx) ? -1 : (a->idx > b->idx);  /* std::sort is unstable: ties are free */
}

/* ShiTomasiFeatureExtractor::extractFeatures (.cpp:5-47): threshold > rmax*quality, raster collect,
 * sort by score descending, first `max`.  Returns the count written. */
ORC_API int orc_shitomasi(const uint8_t *img, int rows, int cols, int step, int max, double quality,
                          int signed_quirk, int *col, int *row, double *score)
{
    size_t n = (size_t)rows * cols;
    double *R = (double *)malloc(sizeof(double) * n);
    orc_shitomasi_response(img, rows, cols, step, signed_quirk, R);
    double rmax = -DBL_MAX;
    for (size_t i = 0; i < n; i++) if (R[i] > rmax) rmax = R[i]; /* minMaxLoc ignores NaN comparisons */
    double thr = rmax * quality;
    dcand_t *c = (dcand_t *)malloc(sizeof(dcand_t) * n);
    int nc = 0;
    for (size_t i = 0; i < n; i++)
        if (R[i] > thr) { c[nc].v = R[i]; c[nc].idx = (int)i; nc++; }
    qsort(c, nc, sizeof(dcand_t), dcand_cmp);
    int out = 0;
    for (int i = 0; i < nc && out < max; i++, out++) {
        row[out] = c[i].idx / cols; col[out] = c[i].idx % cols; score[out] = c[i].v;
    }
    free(c); free(R);
    return out;
}

/* ---- B.4 cv::FAST(img, kp, threshold, nonmax) TYPE_9_16 --------------------------------- */
static const int fast_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int fast_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

static int fast_is_corner(const uint8_t *p, const int *off, int t)
{
    int v = p[0];
    /* >= 9 contiguous circle pixels all brighter than v+t or all darker than v-t */
    for (int sign = 0; sign < 2; sign++) {
        int run = 0;
        for (int k = 0; k < 16 + 8; k++) {
            int q = p[off[k & 15]];
            int ok = sign ? (q < v - t) : (q > v + t);
            if (ok) { if (++run >= 9) return 1; } else run = 0;
        }
    }
    return 0;
}

static int fast_score(const uint8_t *p, const int *off, int threshold)
{
    /* cornerScore<16>: the largest t for which the pixel is still a corner */
    int d[25], v = p[0];
    for (int k = 0; k < 25; k++) d[k] = v - p[off[k & 15]];
    int a0 = threshold;
    for (int k = 0; k < 16; k += 2) {
        int a = d[k + 1] < d[k + 2] ? d[k + 1] : d[k + 2];
        a = a < d[k + 3] ? a : d[k + 3];
        if (a <= a0) continue;
        for (int j = 4; j <= 8; j++) a = a < d[k + j] ? a : d[k + j];
        int t1 = a < d[k] ? a : d[k];
        int t2 = a < d[k + 9] ? a : d[k + 9];
        if (t1 > a0) a0 = t1;
        if (t2 > a0) a0 = t2;
    }
    int b0 = -a0;
    for (int k = 0; k < 16; k += 2) {
        int b = d[k + 1] > d[k + 2] ? d[k + 1] : d[k + 2];
        b = b > d[k + 3] ? b : d[k + 3];
        b = b > d[k + 4] ? b : d[k + 4];
        b = b > d[k + 5] ? b : d[k + 5];
        if (b >= b0) continue;
        for (int j = 6; j <= 8; j++) b = b > d[k + j] ? b : d[k + j];
        int t1 = b > d[k] ? b : d[k];
        int t2 = b > d[k + 9] ? b : d[k + 9];
        if (t1 < b0) b0 = t1;
        if (t2 < b0) b0 = t2;
    }
    return -b0 - 1;
}

/* keypoints in raster order; returns total count (writes at most cap) */
ORC_API int orc_fast(const uint8_t *img, int rows, int cols, int step, int threshold, int nonmax,
                     int *col, int *row, float *score, int cap)
{
    int off[16];
    for (int k = 0; k < 16; k++) off[k] = fast_dy[k] * step + fast_dx[k];
    int *sc = (int *)calloc((size_t)rows * cols, sizeof(int));
    for (int y = 3; y < rows - 3; y++)
        for (int x = 3; x < cols - 3; x++) {
            const uint8_t *p = img + (size_t)y * step + x;
            if (fast_is_corner(p, off, threshold))
                sc[y * cols + x] = nonmax ? fast_score(p, off, threshold) : 1;
        }
    int out = 0;
    for (int y = 3; y < rows - 3; y++)
        for (int x = 3; x < cols - 3; x++) {
            int s = sc[y * cols + x];
            if (!s) continue;
            if (nonmax) {
                int keep = 1;
                for (int j = -1; j <= 1 && keep; j++)
                    for (int i = -1; i <= 1; i++)
                        if ((i || j) && sc[(y + j) * cols + x + i] >= s) { keep = 0; break; }
                if (!keep) continue;
            }
            if (out < cap) { col[out] = x; row[out] = y; score[out] = nonmax ? (float)s : 0.f; }
            out++;
        }
    free(sc);
    return out;
}
