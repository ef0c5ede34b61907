// CPU build of the product's two-view initialiser arithmetic (practical-multi-view_b200/csrc/fivept_math.cuh is
// __host__ __device__) so that tests can pin it to cv2.findEssentialMat / cv2.recoverPose without a GPU.
// Test infrastructure only.
#include "../practical-multi-view_b200/csrc/fivept_math.cuh"
#include <vector>

#define API extern "C" __attribute__((visibility("default")))

// EMEstimatorCallback::runKernel on one 5-point sample (normalised coordinates); returns the number of models
API int fivept_host_models(const double *x1, const double *x2, double *E)
{
    double a[5][2], b[5][2], Em[10][9];
    for (int i = 0; i < 5; i++) { a[i][0] = x1[2 * i]; a[i][1] = x1[2 * i + 1]; b[i][0] = x2[2 * i]; b[i][1] = x2[2 * i + 1]; }
    const int c = fivept::models_from_sample(a, b, Em);
    for (int m = 0; m < c; m++) for (int k = 0; k < 9; k++) E[9 * m + k] = Em[m][k];
    return c;
}

// sequential restatement of cv::findEssentialMat(p1, p2, K, RANSAC, prob, threshold, maxIters, mask): pixel points in,
// E (row major) and the inlier mask out; returns the inlier count of the best model (0: none)
API int fivept_host_find_essential(const double *p1, const double *p2, int n, const double *K, double prob, double threshold,
                                   int max_iters, double *E, unsigned char *mask, int *evaluated)
{
    const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    std::vector<double> x1(2 * n), x2(2 * n);
    for (int i = 0; i < n; i++) {
        x1[2 * i] = (p1[2 * i] - cx) / fx; x1[2 * i + 1] = (p1[2 * i + 1] - cy) / fy;
        x2[2 * i] = (p2[2 * i] - cx) / fx; x2[2 * i + 1] = (p2[2 * i + 1] - cy) / fy;
    }
    threshold /= (fx + fy) / 2;
    const float thr2 = (float)(threshold * threshold);
    pnp::CvRng rng;
    int niters = max_iters > 1 ? max_iters : 1, maxgood = 0, iter = 0;
    for (; iter < niters; iter++) {
        int idx[5];
        for (int i = 0; i < 5;) {
            const int v = rng.uniform(0, n);
            bool dup = false;
            for (int j = 0; j < i; j++) dup = dup || idx[j] == v;
            if (dup) continue;
            idx[i++] = v;
        }
        double a[5][2], b[5][2], Em[10][9];
        for (int i = 0; i < 5; i++) { a[i][0] = x1[2 * idx[i]]; a[i][1] = x1[2 * idx[i] + 1]; b[i][0] = x2[2 * idx[i]]; b[i][1] = x2[2 * idx[i] + 1]; }
        const int nm = fivept::models_from_sample(a, b, Em);
        for (int m = 0; m < nm; m++) {
            int good = 0;
            for (int i = 0; i < n; i++) good += fivept::pair_is_inlier(Em[m], &x1[2 * i], &x2[2 * i], thr2);
            if (good > (maxgood > 4 ? maxgood : 4)) {
                for (int k = 0; k < 9; k++) E[k] = Em[m][k];
                for (int i = 0; i < n; i++) mask[i] = fivept::pair_is_inlier(Em[m], &x1[2 * i], &x2[2 * i], thr2);
                maxgood = good;
                niters = pnp::ransac_update_num_iters(prob, (double)(n - good) / n, 5, niters);
            }
        }
    }
    if (evaluated) *evaluated = iter;
    return maxgood;
}

// cv::recoverPose(E, p1, p2, K, R, t, distanceThresh, mask (in/out), triangulatedPoints (4 x n)); returns the vote count
API int fivept_host_recover_pose(const double *E, const double *p1, const double *p2, int n, const double *K, double dist,
                                 double *R, double *t, unsigned char *mask, double *tri)
{
    const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    double Rc[2][9], tc[3];
    fivept::decompose_essential(E, Rc[0], Rc[1], tc);
    int good[4] = {0, 0, 0, 0};
    std::vector<unsigned char> m(4 * (size_t)n);
    std::vector<double> X(16 * (size_t)n);
    for (int c = 0; c < 4; c++) {
        const double *Rk = Rc[c & 1];
        const double tk[3] = {c < 2 ? tc[0] : -tc[0], c < 2 ? tc[1] : -tc[1], c < 2 ? tc[2] : -tc[2]};
        for (int i = 0; i < n; i++) {
            const double a[2] = {(p1[2 * i] - cx) / fx, (p1[2 * i + 1] - cy) / fy}, b[2] = {(p2[2 * i] - cx) / fx, (p2[2 * i + 1] - cy) / fy};
            double *Xi = &X[((size_t)c * n + i) * 4];
            fivept::triangulate_pair(Rk, tk, a, b, Xi);
            const bool ok = fivept::in_front_of_both(Rk, tk, Xi, dist) && (!mask || mask[i]);
            m[(size_t)c * n + i] = ok;
            good[c] += ok;
        }
    }
    int best;
    if (good[0] >= good[1] && good[0] >= good[2] && good[0] >= good[3]) best = 0;
    else if (good[1] >= good[0] && good[1] >= good[2] && good[1] >= good[3]) best = 1;
    else if (good[2] >= good[0] && good[2] >= good[1] && good[2] >= good[3]) best = 2;
    else best = 3;
    for (int k = 0; k < 9; k++) R[k] = Rc[best & 1][k];
    for (int k = 0; k < 3; k++) t[k] = best < 2 ? tc[k] : -tc[k];
    for (int i = 0; i < n; i++) {
        if (mask) mask[i] = m[(size_t)best * n + i];
        if (tri) for (int k = 0; k < 4; k++) tri[(size_t)k * n + i] = X[((size_t)best * n + i) * 4 + k];
    }
    return good[best];
}
