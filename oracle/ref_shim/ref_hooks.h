// oracle/ref_shim/ref_hooks.h -- TEST INFRASTRUCTURE ONLY.
// Function-pointer hooks through which the shim's cv:: entry points reach the REAL OpenCV kernels of the in-image
// cv2 wheel (tests/bench install them from Python with ctypes callbacks).  A NULL hook falls back to the plain-C
// oracle (oracle/pmv_oracle_*.c, itself pinned to cv2 by tests/test_oracle_*.py), which is linked into the same .so.
#pragma once
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct pmv_ref_hooks {
    /* cv::calcOpticalFlowPyrLK(prev, next, pts, out, status, err, Size(win_w, win_h), max_level) with default criteria / flags */
    int (*lk)(const uint8_t* prev, const uint8_t* next, int rows, int cols, int step_prev, int step_next,
              const float* pts, int n, int win_w, int win_h, int max_level, float* out, uint8_t* status, float* err);
    /* cv::goodFeaturesToTrack(view of parent, max_corners, quality, min_dist, noArray(), 3, 3, false, 0.04) */
    int (*gftt)(const uint8_t* base, int full_rows, int full_cols, int step, int x, int y, int w, int h,
                int max_corners, double quality, double min_dist, int cap, float* xy, int* n_out);
    /* cv::FAST(view, keypoints, threshold, nonmax) */
    int (*fast)(const uint8_t* img, int rows, int cols, int step, int threshold, int nonmax, int cap,
                float* xy, float* response, int* n_out);
    /* cv::blur(src CV_64FC(cn), dst, Size(3,3)) (normalised, BORDER_REFLECT_101), contiguous rows*cols*cn doubles */
    int (*blur3)(const double* src, int rows, int cols, int cn, double* dst);
    /* cv::solvePnPRansac(obj, img, K, noArray(), rvec, tvec, useExtrinsicGuess, iters, reprojErr, confidence, inliers) */
    int (*pnp_ransac)(const float* obj, const float* img, int n, const double* K, double* rvec, double* tvec,
                      int use_guess, int iters, float reproj_err, double confidence, int* inliers, int* n_inliers);
    /* E = cv::findEssentialMat(p1, p2, K, method, prob, threshold, mask); points n x 2 doubles; returns the number of rows of E */
    int (*find_essential)(const double* p1, const double* p2, int n, const double* K, int method, double prob, double threshold,
                          double* E, uint8_t* mask);
    /* cv::recoverPose(E, p1, p2, K, R, t, distanceThresh, mask (in/out), tri (4 x n)); returns the vote count (< 0: failed) */
    int (*recover_pose)(const double* E, const double* p1, const double* p2, int n, const double* K, double distance_thresh,
                        double* R, double* t, uint8_t* mask, double* tri);
} pmv_ref_hooks;
extern pmv_ref_hooks g_pmv_ref_hooks;
#ifdef __cplusplus
}
#endif
