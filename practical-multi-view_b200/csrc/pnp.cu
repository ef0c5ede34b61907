// pnp.cu -- RANSAC pose from 3-D / 2-D correspondences behind BasePnPSolver.
//
// Replaces cv::solvePnPRansac(obj, img, K, noArray(), rvec, tvec, true, 100, 8, .99, inliers) as called by
// OpenCVEPnPSolver::solvePnP (reference OpenCVEPnPSolver.cpp:34-35).  The arithmetic lives in pnp_math.cuh (shared
// with the CPU pin in tests/); this file is the device schedule:
//   * hypotheses are evaluated EIGHT at a time (one warp each: lane 0 runs the 5-point EPnP, the warp then scores all
//     points), the sequential accept rule of RANSACPointSetRegistrator::run -- first hypothesis that beats the best
//     count, loop length shortened by RANSACUpdateNumIters -- is replayed over the batch in order, so the result is
//     the one the sequential loop produces while the latency is that of one hypothesis per batch;
//   * the final Levenberg-Marquardt minimisation over the inliers (solvePnP ITERATIVE from the caller's pose) runs in
//     the same launch, normal equations reduced over the CTA.
// One launch, one small upload (points + subsets), one small download (pose, mask).
#include "common.cuh"
#include "pnp_math.cuh"

namespace {

constexpr int PNP_WARPS = 8;

struct PnpArgs {
    const float *X, *uv;      // n x 3, n x 2
    int n;
    double fu, fv, uc, vc;
    const int *subsets;       // max_iters x 5 (cv::RNG order)
    int max_iters;
    float thr2;
    double conf;
    int use_guess;
    double *pose;             // in: rvec, tvec guess (6) ; out: rvec, tvec (6)
    unsigned char *mask_best; // n
    unsigned char *mask_tmp;  // PNP_WARPS x n
    int *info;                // [0] inliers of the best hypothesis (0: none found), [1] hypotheses evaluated, [2] LM iterations
};

__global__ void __launch_bounds__(PNP_WARPS * 32) pnp_ransac_kernel(const PnpArgs A)
{
    __shared__ double Rh[PNP_WARPS][9], th[PNP_WARPS][3];
    __shared__ int cnt[PNP_WARPS];
    __shared__ double Rb[9], tb[3];          // best hypothesis, then the running LM pose
    __shared__ double Rc[9], tc[3];          // LM candidate
    __shared__ double red[PNP_WARPS][28];
    __shared__ int s_niters, s_maxgood, s_take, s_flag;
    __shared__ double s_H[21], s_g[6], s_cost, s_cost2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = A.n;
    if (tid == 0) { s_niters = A.max_iters; s_maxgood = 0; }
    __syncthreads();
    int evaluated = 0;
    for (int base = 0; base < s_niters; base += PNP_WARPS) {
        const int h = base + warp;
        const bool live = h < s_niters;      // s_niters only changes between the barriers below
        if (live && lane == 0) {
            pnp::EPnP e;
            e.n = 5; e.fu = A.fu; e.fv = A.fv; e.uc = A.uc; e.vc = A.vc;
            for (int i = 0; i < 5; i++) {
                const int p = A.subsets[5 * h + i];
                for (int k = 0; k < 3; k++) e.pws[i][k] = A.X[3 * p + k];
                e.us[i][0] = A.uv[2 * p]; e.us[i][1] = A.uv[2 * p + 1];
            }
            double R[9], t[3], r[3];
            e.compute_pose(R, t);
            pnp::matrix_to_rodrigues(R, r);      // the hypothesis travels as [rvec | tvec]
            pnp::rodrigues_to_matrix(r, R);
            for (int k = 0; k < 9; k++) Rh[warp][k] = R[k];
            for (int k = 0; k < 3; k++) th[warp][k] = t[k];
        }
        __syncwarp();
        int c = 0;
        if (live) {
            unsigned char *m = A.mask_tmp + (size_t)warp * n;
            for (int i = lane; i < n; i += 32) {
                const bool in = pnp::point_is_inlier(A.X + 3 * i, A.uv + 2 * i, Rh[warp], th[warp], A.fu, A.fv, A.uc, A.vc, A.thr2);
                m[i] = in;
                c += in;
            }
            for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        }
        if (lane == 0) cnt[warp] = live ? c : -1;
        __syncthreads();
        // the sequential accept rule, hypothesis by hypothesis
        for (int w = 0; w < PNP_WARPS; w++) {
            if (tid == 0) {
                s_take = 0;
                if (base + w < s_niters && cnt[w] >= 0) {
                    evaluated = base + w + 1;
                    if (cnt[w] > (s_maxgood > 4 ? s_maxgood : 4)) {
                        s_take = 1;
                        s_maxgood = cnt[w];
                        for (int k = 0; k < 9; k++) Rb[k] = Rh[w][k];
                        for (int k = 0; k < 3; k++) tb[k] = th[w][k];
                        s_niters = pnp::ransac_update_num_iters(A.conf, (double)(n - cnt[w]) / n, 5, s_niters);
                    }
                }
            }
            __syncthreads();
            if (s_take) {
                const unsigned char *m = A.mask_tmp + (size_t)w * n;
                for (int i = tid; i < n; i += PNP_WARPS * 32) A.mask_best[i] = m[i];
            }
            __syncthreads();
        }
    }
    if (tid == 0) { A.info[0] = s_maxgood; A.info[1] = evaluated; A.info[2] = 0; }
    if (s_maxgood == 0) return;              // no model: the pose stays the caller's (cv::solvePnPRansac returns false)
    // ---- refinement over the inliers from the caller's pose (useExtrinsicGuess) or from the best hypothesis
    if (tid == 0) {
        if (A.use_guess) {
            pnp::rodrigues_to_matrix(A.pose, Rb);
            for (int k = 0; k < 3; k++) tb[k] = A.pose[3 + k];
        }
        s_flag = 0;
    }
    __syncthreads();
    auto accumulate = [&](const double *R, const double *t, bool with_jac) {
        double acc[28];
#pragma unroll
        for (int k = 0; k < 28; k++) acc[k] = 0.0;
        for (int i = tid; i < n; i += PNP_WARPS * 32) {
            if (!A.mask_best[i]) continue;
            double r[2], J[12];
            pnp::reproj_jac(A.X + 3 * i, A.uv + 2 * i, R, t, A.fu, A.fv, A.uc, A.vc, r, J);
            if (with_jac) pnp::lm_accumulate(r, J, acc, acc + 21, acc + 27);
            else acc[27] += r[0] * r[0] + r[1] * r[1];
        }
        const int k0 = with_jac ? 0 : 27;
        for (int k = k0; k < 28; k++) {
            double v = acc[k];
            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[warp][k] = v;
        }
        __syncthreads();
    };
    double lambda = 1e-3;
    int it = 0;
    for (; it < 100; it++) {
        accumulate(Rb, tb, true);
        if (tid == 0) {
            for (int k = 0; k < 21; k++) { double v = 0; for (int w = 0; w < PNP_WARPS; w++) v += red[w][k]; s_H[k] = v; }
            for (int k = 0; k < 6; k++) { double v = 0; for (int w = 0; w < PNP_WARPS; w++) v += red[w][21 + k]; s_g[k] = v; }
            double v = 0; for (int w = 0; w < PNP_WARPS; w++) v += red[w][27];
            s_cost = v;
        }
        __syncthreads();
        bool moved = false;
        double dn = 0.0;
        for (int tries = 0; tries < 30 && !moved; tries++) {
            if (tid == 0) {
                double d[6];
                s_flag = pnp::lm_solve(s_H, s_g, lambda, d) ? 1 : 0;
                if (s_flag) {
                    pnp::lm_apply(Rb, tb, d, Rc, tc);
                    s_cost2 = 0; for (int k = 0; k < 6; k++) s_cost2 += d[k] * d[k];   // |step|^2, parked here until the cost comes back
                }
            }
            __syncthreads();
            if (!s_flag) { lambda *= 10; __syncthreads(); continue; }
            const double step2 = s_cost2;
            __syncthreads();
            accumulate(Rc, tc, false);
            double c2 = 0;
            for (int w = 0; w < PNP_WARPS; w++) c2 += red[w][27];
            __syncthreads();
            if (c2 <= s_cost) {
                if (tid == 0) {
                    for (int k = 0; k < 9; k++) Rb[k] = Rc[k];
                    for (int k = 0; k < 3; k++) tb[k] = tc[k];
                }
                dn = step2;
                lambda = lambda > 1e-12 ? lambda * 0.1 : lambda;
                moved = true;
            } else {
                lambda *= 10;
            }
            __syncthreads();
        }
        if (!moved || dn < 1e-24) break;
    }
    if (tid == 0) {
        double r[3];
        pnp::matrix_to_rodrigues(Rb, r);
        for (int k = 0; k < 3; k++) { A.pose[k] = r[k]; A.pose[3 + k] = tb[k]; }
        A.info[2] = it;
    }
}

}  // namespace

extern "C" {

PMV_API int pmv_pnp_ransac(pmv_ctx *ctx, const float *obj_xyz, const float *img_xy, int n, const double K[9], double rvec[3],
                           double tvec[3], int use_extrinsic_guess, int iterations, float reproj_err, double confidence,
                           uint8_t *inlier_mask, int *n_inliers)
{
    if (!ctx) return PMV_ERR_INVALID;
    if (!obj_xyz || !img_xy || !K || !rvec || !tvec || !n_inliers || iterations <= 0 || reproj_err <= 0 || n < 0)
        return ctx->fail(PMV_ERR_INVALID, "pmv_pnp_ransac: bad argument");
    if (n < 6)   // OpenCV switches to a direct EPnP / P3P solve for n == 5 / 4 and throws below 4
        return ctx->fail(PMV_ERR_UNSUPPORTED, "pmv_pnp_ransac: fewer than 6 correspondences");
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    // subsets in cv::RNG((uint64)-1) order: the registrator draws every subset of the loop whether or not it is used
    const size_t b_pts = (size_t)n * 5 * sizeof(float), b_sub = (size_t)iterations * 5 * sizeof(int);
    const size_t off_sub = (b_pts + 15) & ~(size_t)15, off_pose = (off_sub + b_sub + 15) & ~(size_t)15;
    const size_t off_info = off_pose + 64, off_mask = off_info + 64, off_tmp = (off_mask + n + 15) & ~(size_t)15;
    const size_t total = off_tmp + (size_t)PNP_WARPS * n + 16;
    cudaError_t e = ctx->scratch[0].reserve(total);
    if (e == cudaSuccess) e = ctx->pin[1].reserve(total);
    if (e != cudaSuccess) return ctx->fail(PMV_ERR_NOMEM, "pnp workspace", e);
    char *h = ctx->pin[1].as<char>(), *d = ctx->scratch[0].as<char>();
    memcpy(h, obj_xyz, (size_t)n * 3 * sizeof(float));
    memcpy(h + (size_t)n * 3 * sizeof(float), img_xy, (size_t)n * 2 * sizeof(float));
    {
        pnp::CvRng rng;
        int *sub = reinterpret_cast<int *>(h + off_sub);
        for (int it = 0; it < iterations; it++) {
            int *idx = sub + 5 * it;
            for (int i = 0; i < 5;) {
                const int v = rng.uniform(0, n);
                bool dup = false;
                for (int j = 0; j < i; j++) dup = dup || idx[j] == v;
                if (!dup) idx[i++] = v;
            }
        }
    }
    double *hp = reinterpret_cast<double *>(h + off_pose);
    for (int k = 0; k < 3; k++) { hp[k] = rvec[k]; hp[3 + k] = tvec[k]; }
    PMV_CUDA_TRY(ctx, cudaMemcpyAsync(d, h, off_pose + 64, cudaMemcpyHostToDevice, s));
    PnpArgs A;
    A.X = reinterpret_cast<const float *>(d); A.uv = A.X + 3 * (size_t)n; A.n = n;
    A.fu = K[0]; A.fv = K[4]; A.uc = K[2]; A.vc = K[5];
    A.subsets = reinterpret_cast<const int *>(d + off_sub); A.max_iters = iterations;
    A.thr2 = reproj_err * reproj_err; A.conf = confidence; A.use_guess = use_extrinsic_guess;
    A.pose = reinterpret_cast<double *>(d + off_pose);
    A.info = reinterpret_cast<int *>(d + off_info);
    A.mask_best = reinterpret_cast<unsigned char *>(d + off_mask);
    A.mask_tmp = reinterpret_cast<unsigned char *>(d + off_tmp);
    pnp_ransac_kernel<<<1, PNP_WARPS * 32, 0, s>>>(A);
    PMV_LAUNCH_CHECK(ctx, "pnp_ransac_kernel");
    PMV_CUDA_TRY(ctx, cudaMemcpyAsync(h + off_pose, d + off_pose, 128 + (size_t)n, cudaMemcpyDeviceToHost, s));
    PMV_CUDA_TRY(ctx, cudaStreamSynchronize(s));
    const int *info = reinterpret_cast<const int *>(h + off_info);
    *n_inliers = info[0];
    if (info[0] > 0) {
        for (int k = 0; k < 3; k++) { rvec[k] = hp[k]; tvec[k] = hp[3 + k]; }
        if (inlier_mask) memcpy(inlier_mask, h + off_mask, n);
    } else if (inlier_mask) {
        memset(inlier_mask, 0, n);
    }
    return PMV_OK;
}

}  // extern "C"
