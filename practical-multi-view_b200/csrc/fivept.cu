// fivept.cu -- two-view initialisation behind BaseTriangulator.
//
// Replaces, for OpenCVFivePointTri::triangulate (reference OpenCVFivePointTri.cpp:25-27),
//   E = cv::findEssentialMat(p1, p2, camera, cv::RANSAC, 0.99, 1, mask);
//   cv::recoverPose(E, p1, p2, camera, R_out, t_out, HUGE_VAL, mask, tri);
// The arithmetic lives in fivept_math.cuh (shared with the CPU pin in tests/); this file is the device schedule, ONE
// launch of one CTA for both calls:
//   * RANSAC: samples are drawn in cv::RNG((uint64)-1) order by thread 0 and solved 256 at a time, one per THREAD (the
//     five-point solver is a long serial fp64 chain: up to ten essential matrices per sample); each warp then scores the
//     models of its 32 samples over all correspondences (Sampson distance, float compare); the sequential accept rule of
//     RANSACPointSetRegistrator::run -- first model that beats the best count, loop length shortened by
//     RANSACUpdateNumIters -- is then replayed over the batch in (sample, model) order, so the result is the one the
//     sequential loop produces (samples beyond the shortened loop are simply ignored);
//   * recoverPose: E is decomposed once, every thread triangulates its correspondences under the four pose candidates
//     (4 x 4 Jacobi SVD each), the cheirality votes are reduced over the CTA, and the winner's pose / mask / points are
//     written out.
// One small upload (points), one small download (E, R, t, masks, points).
#include "common.cuh"
#include "fivept_math.cuh"

namespace {

constexpr int FP_WARPS = 8;
constexpr int FP_THREADS = FP_WARPS * 32;

struct FivePtArgs {
    const double *p1, *p2;    // n x 2 pixel coordinates
    double *x1, *x2;          // n x 2 normalised coordinates
    double *X4;               // 4 candidates x n x 4 homogeneous points
    double *models;           // FP_THREADS samples x 10 models x 9
    double *out;              // E (9) | R (9) | t (3); E is an input when !do_ransac
    int *info;                // [0] RANSAC inliers (0: no model), [1] samples drawn, [2] votes of the winner, [3] winner (0..3)
    unsigned char *mask_ransac, *mask, *mask_cand;   // n | n (in: optional caller mask, out: final mask) | 4 x n
    int n;
    double fx, fy, cx, cy;
    double prob, thr, dist;
    int max_iters, do_ransac, do_recover, has_mask;
};

__global__ void __launch_bounds__(FP_THREADS) fivept_pose_kernel(const FivePtArgs A)
{
    __shared__ int nm[FP_THREADS], cnt[FP_THREADS][10], sub[FP_THREADS][5];
    __shared__ double Eb[9], Rc[2][9], tc[3];
    __shared__ int s_niters, s_maxgood, s_drawn, votes[4], s_best;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = A.n;
    for (int i = tid; i < n; i += FP_THREADS) {
        A.x1[2 * i] = (A.p1[2 * i] - A.cx) / A.fx; A.x1[2 * i + 1] = (A.p1[2 * i + 1] - A.cy) / A.fy;
        A.x2[2 * i] = (A.p2[2 * i] - A.cx) / A.fx; A.x2[2 * i + 1] = (A.p2[2 * i + 1] - A.cy) / A.fy;
    }
    if (tid == 0) { s_niters = A.max_iters > 1 ? A.max_iters : 1; s_maxgood = 0; s_drawn = 0; }
    if (tid < 9) Eb[tid] = A.do_ransac ? 0.0 : A.out[tid];
    __syncthreads();
    if (A.do_ransac) {
        const double thr = A.thr / ((A.fx + A.fy) / 2);
        const float thr2 = (float)(thr * thr);
        pnp::CvRng rng;                       // thread 0's copy is the one that advances
        for (int base = 0; base < s_niters; base += FP_THREADS) {
            if (tid == 0) {
                for (int w = 0; w < FP_THREADS; w++)
                    for (int i = 0; i < 5;) {
                        const int v = rng.uniform(0, n);
                        bool dup = false;
                        for (int j = 0; j < i; j++) dup = dup || sub[w][j] == v;
                        if (!dup) sub[w][i++] = v;
                    }
            }
            __syncthreads();
            const int niters = s_niters;            // only changes between the barriers below
            // one sample per THREAD: the five-point solver is a long serial fp64 chain, 256 of them run side by side
            if (base + tid < niters) {
                double a[5][2], b[5][2];
                for (int i = 0; i < 5; i++) {
                    const int p = sub[tid][i];
                    a[i][0] = A.x1[2 * p]; a[i][1] = A.x1[2 * p + 1];
                    b[i][0] = A.x2[2 * p]; b[i][1] = A.x2[2 * p + 1];
                }
                nm[tid] = fivept::models_from_sample(a, b, reinterpret_cast<double (*)[9]>(A.models + (size_t)tid * 90));
            } else {
                nm[tid] = 0;
            }
            __syncthreads();
            // each warp scores the models of its 32 samples over all correspondences
            for (int k = 0; k < 32; k++) {
                const int sm = warp * 32 + k, c = nm[sm];
                for (int m = 0; m < c; m++) {
                    const double *Em = A.models + ((size_t)sm * 10 + m) * 9;
                    int good = 0;
                    for (int i = lane; i < n; i += 32) good += fivept::pair_is_inlier(Em, A.x1 + 2 * i, A.x2 + 2 * i, thr2);
                    for (int o = 16; o; o >>= 1) good += __shfl_xor_sync(0xffffffffu, good, o);
                    if (lane == 0) cnt[sm][m] = good;
                }
            }
            __syncthreads();
            if (tid == 0) {
                for (int w = 0; w < FP_THREADS && base + w < s_niters; w++) {
                    s_drawn = base + w + 1;
                    for (int m = 0; m < nm[w]; m++)
                        if (cnt[w][m] > (s_maxgood > 4 ? s_maxgood : 4)) {
                            s_maxgood = cnt[w][m];
                            for (int k = 0; k < 9; k++) Eb[k] = A.models[((size_t)w * 10 + m) * 9 + k];
                            s_niters = pnp::ransac_update_num_iters(A.prob, (double)(n - cnt[w][m]) / n, 5, s_niters);
                        }
                }
            }
            __syncthreads();
        }
        for (int i = tid; i < n; i += FP_THREADS)
            A.mask_ransac[i] = s_maxgood > 0 && fivept::pair_is_inlier(Eb, A.x1 + 2 * i, A.x2 + 2 * i, thr2);
        if (tid < 9) A.out[tid] = Eb[tid];
        if (tid == 0) { A.info[0] = s_maxgood; A.info[1] = s_drawn; A.info[2] = 0; A.info[3] = 0; }
        if (s_maxgood == 0 || !A.do_recover) return;
    }
    // ---- cv::recoverPose
    if (tid == 0) fivept::decompose_essential(Eb, Rc[0], Rc[1], tc);
    if (tid < 4) votes[tid] = 0;
    __syncthreads();
    int v[4] = {0, 0, 0, 0};
    for (int i = tid; i < n; i += FP_THREADS) {
        const bool keep = A.do_ransac ? A.mask_ransac[i] != 0 : (A.has_mask ? A.mask[i] != 0 : true);
        for (int c = 0; c < 4; c++) {
            const double *Rk = Rc[c & 1];
            const double sg = c < 2 ? 1.0 : -1.0;
            const double tk[3] = {sg * tc[0], sg * tc[1], sg * tc[2]};
            double X[4];
            fivept::triangulate_pair(Rk, tk, A.x1 + 2 * i, A.x2 + 2 * i, X);
            const bool ok = fivept::in_front_of_both(Rk, tk, X, A.dist) && keep;
            A.mask_cand[(size_t)c * n + i] = ok;
            v[c] += ok;
            double *Xo = A.X4 + ((size_t)c * n + i) * 4;
            for (int k = 0; k < 4; k++) Xo[k] = X[k];
        }
    }
    for (int c = 0; c < 4; c++) {
        int s = v[c];
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0 && s) atomicAdd(&votes[c], s);
    }
    __syncthreads();
    if (tid == 0) {
        const int g0 = votes[0], g1 = votes[1], g2 = votes[2], g3 = votes[3];
        int best;
        if (g0 >= g1 && g0 >= g2 && g0 >= g3) best = 0;
        else if (g1 >= g0 && g1 >= g2 && g1 >= g3) best = 1;
        else if (g2 >= g0 && g2 >= g1 && g2 >= g3) best = 2;
        else best = 3;
        s_best = best;
        for (int k = 0; k < 9; k++) A.out[9 + k] = Rc[best & 1][k];
        for (int k = 0; k < 3; k++) A.out[18 + k] = best < 2 ? tc[k] : -tc[k];
        A.info[2] = votes[best]; A.info[3] = best;
    }
    __syncthreads();
    const int best = s_best;
    for (int i = tid; i < n; i += FP_THREADS) A.mask[i] = A.mask_cand[(size_t)best * n + i];
    // the winner's points, 4 x n like cv::triangulatePoints returns them
    for (int i = tid; i < 4 * n; i += FP_THREADS) {
        const int k = i / n, p = i - k * n;
        A.X4[(size_t)16 * n + i] = A.X4[((size_t)best * n + p) * 4 + k];
    }
}

int run(pmv_ctx *ctx, const double *p1, const double *p2, int n, const double K[9], double prob, double threshold, int max_iters,
        double dist, int do_ransac, int do_recover, double *E, double *R, double *t, uint8_t *ransac_mask, uint8_t *mask,
        double *tri, int *n_inliers, int *n_good)
{
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    const size_t nn = (size_t)n;
    // doubles: p1 p2 | x1 x2 | X4 (16 n) + winner (4 n) | out (24, padded to 32) | models
    const size_t off_x = 4 * nn, off_X4 = 8 * nn, off_out = 28 * nn, off_models = off_out + 32, n_dbl = off_models + (size_t)FP_THREADS * 90;
    const size_t off_info = n_dbl * sizeof(double), off_mr = off_info + 64, off_m = (off_mr + nn + 15) & ~(size_t)15,
                 off_mc = (off_m + nn + 15) & ~(size_t)15, total = off_mc + 4 * nn + 16;
    cudaError_t e = ctx->scratch[0].reserve(total);
    if (e == cudaSuccess) e = ctx->pin[1].reserve(total);
    if (e != cudaSuccess) return ctx->fail(PMV_ERR_NOMEM, "five-point workspace", e);
    char *h = ctx->pin[1].as<char>(), *d = ctx->scratch[0].as<char>();
    double *hd = reinterpret_cast<double *>(h), *dd = reinterpret_cast<double *>(d);
    memcpy(hd, p1, 2 * nn * sizeof(double));
    memcpy(hd + 2 * nn, p2, 2 * nn * sizeof(double));
    PMV_CUDA_TRY(ctx, cudaMemcpyAsync(d, h, 4 * nn * sizeof(double), cudaMemcpyHostToDevice, s));
    const int has_mask = !do_ransac && mask != nullptr;
    if (!do_ransac) {
        for (int k = 0; k < 9; k++) hd[off_out + k] = E[k];
        PMV_CUDA_TRY(ctx, cudaMemcpyAsync(dd + off_out, hd + off_out, 9 * sizeof(double), cudaMemcpyHostToDevice, s));
        if (has_mask) {
            memcpy(h + off_m, mask, nn);
            PMV_CUDA_TRY(ctx, cudaMemcpyAsync(d + off_m, h + off_m, nn, cudaMemcpyHostToDevice, s));
        }
    }
    FivePtArgs A;
    A.p1 = dd; A.p2 = dd + 2 * nn; A.x1 = dd + off_x; A.x2 = dd + off_x + 2 * nn; A.X4 = dd + off_X4; A.out = dd + off_out; A.models = dd + off_models;
    A.info = reinterpret_cast<int *>(d + off_info);
    A.mask_ransac = reinterpret_cast<unsigned char *>(d + off_mr);
    A.mask = reinterpret_cast<unsigned char *>(d + off_m);
    A.mask_cand = reinterpret_cast<unsigned char *>(d + off_mc);
    A.n = n; A.fx = K[0]; A.fy = K[4]; A.cx = K[2]; A.cy = K[5];
    A.prob = prob; A.thr = threshold; A.dist = dist; A.max_iters = max_iters;
    A.do_ransac = do_ransac; A.do_recover = do_recover; A.has_mask = has_mask;
    PMV_CUDA_TRY(ctx, cudaMemsetAsync(d + off_info, 0, 64, s));
    fivept_pose_kernel<<<1, FP_THREADS, 0, s>>>(A);
    PMV_LAUNCH_CHECK(ctx, "fivept_pose_kernel");
    // winner's points + out + info + masks come back in two copies
    PMV_CUDA_TRY(ctx, cudaMemcpyAsync(hd + off_X4 + 16 * nn, dd + off_X4 + 16 * nn, (4 * nn + 32) * sizeof(double), cudaMemcpyDeviceToHost, s));
    PMV_CUDA_TRY(ctx, cudaMemcpyAsync(h + off_info, d + off_info, off_mc - off_info, cudaMemcpyDeviceToHost, s));
    PMV_CUDA_TRY(ctx, cudaStreamSynchronize(s));
    const int *info = reinterpret_cast<const int *>(h + off_info);
    if (do_ransac) {
        if (n_inliers) *n_inliers = info[0];
        for (int k = 0; k < 9; k++) E[k] = hd[off_out + k];
        if (ransac_mask) memcpy(ransac_mask, h + off_mr, nn);
    }
    const bool recovered = do_recover && (!do_ransac || info[0] > 0);
    if (n_good) *n_good = recovered ? info[2] : 0;
    if (recovered) {
        if (R) for (int k = 0; k < 9; k++) R[k] = hd[off_out + 9 + k];
        if (t) for (int k = 0; k < 3; k++) t[k] = hd[off_out + 18 + k];
        if (mask) memcpy(mask, h + off_m, nn);
        if (tri) memcpy(tri, hd + off_X4 + 16 * nn, 4 * nn * sizeof(double));
    } else if (do_recover) {
        if (mask) memset(mask, 0, nn);
    }
    return PMV_OK;
}

}  // namespace

extern "C" {

PMV_API int pmv_find_essential_mat(pmv_ctx *ctx, const double *p1_xy, const double *p2_xy, int n, const double K[9], double prob,
                                   double threshold, int max_iters, double E[9], uint8_t *mask, int *n_inliers)
{
    if (!ctx) return PMV_ERR_INVALID;
    if (!p1_xy || !p2_xy || !K || !E || !n_inliers || threshold <= 0 || prob <= 0 || prob >= 1)
        return ctx->fail(PMV_ERR_INVALID, "pmv_find_essential_mat: bad argument");
    if (n < 6)   // five points return every root stacked, fewer throw (cv::findEssentialMat)
        return ctx->fail(PMV_ERR_UNSUPPORTED, "pmv_find_essential_mat: fewer than 6 correspondences");
    return run(ctx, p1_xy, p2_xy, n, K, prob, threshold, max_iters, 0.0, 1, 0, E, nullptr, nullptr, mask, nullptr, nullptr, n_inliers, nullptr);
}

PMV_API int pmv_recover_pose(pmv_ctx *ctx, const double E[9], const double *p1_xy, const double *p2_xy, int n, const double K[9],
                             double distance_thresh, double R[9], double t[3], uint8_t *mask, double *tri, int *n_good)
{
    if (!ctx) return PMV_ERR_INVALID;
    if (!E || !p1_xy || !p2_xy || !K || !R || !t || n < 1) return ctx->fail(PMV_ERR_INVALID, "pmv_recover_pose: bad argument");
    double Ec[9];
    for (int k = 0; k < 9; k++) Ec[k] = E[k];
    return run(ctx, p1_xy, p2_xy, n, K, 0.0, 0.0, 0, distance_thresh, 0, 1, Ec, R, t, nullptr, mask, tri, nullptr, n_good);
}

PMV_API int pmv_five_point_pose(pmv_ctx *ctx, const double *p1_xy, const double *p2_xy, int n, const double K[9], double prob,
                                double threshold, int max_iters, double distance_thresh, double E[9], double R[9], double t[3],
                                uint8_t *ransac_mask, uint8_t *mask, double *tri, int *n_inliers, int *n_good)
{
    if (!ctx) return PMV_ERR_INVALID;
    if (!p1_xy || !p2_xy || !K || !E || !R || !t || !n_inliers || threshold <= 0 || prob <= 0 || prob >= 1)
        return ctx->fail(PMV_ERR_INVALID, "pmv_five_point_pose: bad argument");
    if (n < 6) return ctx->fail(PMV_ERR_UNSUPPORTED, "pmv_five_point_pose: fewer than 6 correspondences");
    return run(ctx, p1_xy, p2_xy, n, K, prob, threshold, max_iters, distance_thresh, 1, 1, E, R, t, ransac_mask, mask, tri, n_inliers, n_good);
}

}  // extern "C"
