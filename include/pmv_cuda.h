/*
 * pmv_cuda.h -- C ABI of libpmv_cuda.so: the B200 (sm_100a) replacement for the
 * visual-odometry hot path of JeanElsner/practical-multi-view.
 *
 * Every entry point below replaces one call the reference's plugin layer makes into
 * OpenCV / Ceres (reference file:line cited per function; paths relative to the
 * reference checkout).  The header is plain C: POD arguments, caller-owned buffers,
 * an opaque context that owns device memory and one CUDA stream.  No exception crosses
 * the ABI; every function returns PMV_OK (0) or a negative pmv_status, and
 * pmv_last_error() gives the text.  There is NO CPU fallback: without a CUDA device
 * pmv_create() returns NULL and every other call fails.
 *
 * Threading (OdometryPipeline.cpp:210-245): the matcher and extractor run on the
 * producer thread, the optimizer on the consumer thread -> contexts are independent
 * and re-entrant across handles; one context must not be used from two threads at once.
 *
 * "host" entry points take host pointers (pageable or pinned) and include the H2D/D2H
 * copies; "_dev" entry points take device pointers resident in HBM and enqueue on the
 * context's stream without synchronising.
 */
#ifndef PMV_CUDA_H
#define PMV_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PMV_API __attribute__((visibility("default")))
#else
#define PMV_API
#endif

typedef struct pmv_ctx pmv_ctx;

typedef enum pmv_status {
    PMV_OK = 0,
    PMV_ERR_INVALID = -1,     /* bad argument (null pointer, non-positive size, ...) */
    PMV_ERR_UNSUPPORTED = -2, /* valid for the reference but outside what the kernels cover */
    PMV_ERR_CUDA = -3,        /* CUDA runtime / launch failure, see pmv_last_error */
    PMV_ERR_NOMEM = -4,
    PMV_ERR_NCCL = -5,
    PMV_ERR_NUMERIC = -6      /* e.g. reduced camera system not positive definite */
} pmv_status;

#define PMV_MAX_PYR_LEVELS 8 /* level 0 + up to 7 reduced levels */

/* flags of pmv_lk_track == cv::OPTFLOW_* (video/tracking.hpp) */
#define PMV_LK_USE_INITIAL_FLOW 4
#define PMV_LK_GET_MIN_EIGENVALS 8

/* ------------------------------------------------------------------ context ---------- */
/* One per adapter object (GpuLucasKanadeFM / Gpu*FeatureExtractor / GpuBundleAdjustment own
 * one each; replaces nothing in the reference -- it has no device state).  NULL on failure. */
PMV_API pmv_ctx *pmv_create(int device);
PMV_API void pmv_destroy(pmv_ctx *ctx);
/* Adopt a caller-owned cudaStream_t (e.g. torch's current stream); NULL = context's own. */
PMV_API int pmv_set_stream(pmv_ctx *ctx, void *cuda_stream);
PMV_API int pmv_sync(pmv_ctx *ctx);
PMV_API const char *pmv_last_error(pmv_ctx *ctx);
/* Number of kernel launches this context has issued since creation (bench "gpu_launches"). */
PMV_API uint64_t pmv_launch_count(pmv_ctx *ctx);
PMV_API const char *pmv_version(void);

/* Per-phase device timing for bench.py's roofline: when enabled every API call brackets its
 * kernel groups with CUDA events on the context stream.  pmv_profile_collect synchronises,
 * returns the summed milliseconds and the number of bracketed groups per phase since the last
 * collect, and resets. */
#define PMV_PHASE_PYRAMID 0
#define PMV_PHASE_LK 1
#define PMV_PHASE_RESPONSE 2
#define PMV_PHASE_SELECT 3
#define PMV_PHASE_FAST 4
#define PMV_PHASE_BA 5
#define PMV_PHASE_PYR_L0 6   /* the level-0 launch of the pyramid group (inside PMV_PHASE_PYRAMID) */
#define PMV_PHASE_COUNT 8
PMV_API int pmv_profile_enable(pmv_ctx *ctx, int on);
PMV_API int pmv_profile_collect(pmv_ctx *ctx, int n_phases, double *ms_sum, int *count);

/* Measured fp64 peaks of this context's device ("fp64 ALU peaks must be measured on the box"): a DFMA
 * chain and an mma.sync.m8n8k4.f64 (DMMA) chain, TFLOP/s.  Diagnostic for bench.py's fp64 rooflines (the window bundle
 * adjuster, CeresBundleAdjustment.cpp:54-61, is fp64-bound); replaces nothing in the reference. */
PMV_API int pmv_probe_fp64(pmv_ctx *ctx, double *dfma_tflops, double *dmma_tflops);

/* ------------------------------------------------------------------ pyramid ---------- */
/* Effective top level of cv::buildOpticalFlowPyramid(img, Size(win_w,win_h), max_level):
 * max_level is clipped when the next level would be <= the window (SURVEY Appx A.1).
 * Pure host arithmetic.  Reached from OpenCVLucasKanadeFM.cpp:15. */
PMV_API int pmv_pyr_levels(int rows, int cols, int win_w, int win_h, int max_level);

/* Gaussian 5x5 pyramid (cv::pyrDown chain inside calcOpticalFlowPyrLK,
 * OpenCVLucasKanadeFM.cpp:15).  Host image in, levels 1..L out, packed back to back
 * (level l is rows_l x cols_l, no padding).  *out_levels = L. Bit-exact vs cv::pyrDown. */
PMV_API int pmv_pyramid_build(pmv_ctx *ctx, const uint8_t *img, int rows, int cols, int step,
                              int win_w, int win_h, int max_level,
                              uint8_t *out_packed, size_t out_capacity, int *out_levels);

/* Scharr derivative image, int16 x2 interleaved [Ix,Iy] (cv::calcSharrDeriv inside
 * calcOpticalFlowPyrLK).  The LK kernel fuses this; the entry point exists for
 * stage-by-stage parity.  out = rows*cols*2 int16. */
PMV_API int pmv_scharr(pmv_ctx *ctx, const uint8_t *img, int rows, int cols, int step, int16_t *out);

/* ------------------------------------------------------------------ Lucas-Kanade ------ */
/* == cv::calcOpticalFlowPyrLK(prev, next, prevPts, nextPts, status, err, Size(win_w,win_h),
 *    max_level, TermCriteria(COUNT+EPS, max_count, eps), flags, min_eig_thr)
 * as called at OpenCVLucasKanadeFM.cpp:15 (win 32x32, max_level 4, defaults 30 / 0.01 / 0 / 1e-4).
 * prev_xy/next_xy: n x (x,y) float; status: n bytes; err: n floats.  With
 * PMV_LK_USE_INITIAL_FLOW next_xy is also an input.  Windows up to 32x32 (w*h <= 1024). */
PMV_API int pmv_lk_track(pmv_ctx *ctx, const uint8_t *prev, const uint8_t *next,
                         int rows, int cols, int step, const float *prev_xy, int n,
                         int win_w, int win_h, int max_level, int max_count, double eps,
                         int flags, double min_eig_thr,
                         float *next_xy, uint8_t *status, float *err);

/* B independent frame pairs in one call (BASELINE config 2).  Image b of prev/next starts at
 * b*img_stride bytes; points of pair b at prev_xy + b*n*2.  Host buffers; uploads are chunked
 * and overlapped with the kernels. */
PMV_API int pmv_lk_track_batched(pmv_ctx *ctx, const uint8_t *prev, const uint8_t *next, int batch,
                                 size_t img_stride, int rows, int cols, int step,
                                 const float *prev_xy, int n, int win_w, int win_h, int max_level,
                                 int max_count, double eps, int flags, double min_eig_thr,
                                 float *next_xy, uint8_t *status, float *err);

/* Same, everything resident in HBM (device pointers); asynchronous on the context stream. */
PMV_API int pmv_lk_track_batched_dev(pmv_ctx *ctx, const uint8_t *d_prev, const uint8_t *d_next,
                                     int batch, size_t img_stride, int rows, int cols, int step,
                                     const float *d_prev_xy, int n, int win_w, int win_h,
                                     int max_level, int max_count, double eps, int flags,
                                     double min_eig_thr,
                                     float *d_next_xy, uint8_t *d_status, float *d_err);

/* ------------------------------------------------------------------ resident front end -- */
/* One handle for the front end of OdometryPipeline::addFrame (OdometryPipeline.cpp:329-374) with the images, both
 * pyramids, the Scharr planes and the track list RESIDENT on the device between frames.  Per frame: one upload, one
 * fused pyramid build (the reference reduces every frame twice, OpenCVLucasKanadeFM.cpp:15), the LK solve from the
 * resident features of the previous frame, the status filter + Feature(int, int) truncation
 * (OpenCVLucasKanadeFM.cpp:23-29), and -- when fewer than tracked_tol survive -- the ROI-grid goodFeaturesToTrack
 * re-extraction on the PREVIOUS frame (getGridROI, OdometryPipeline.cpp:351,674-692) with Frame::hasNeighbor
 * de-duplication (Frame.cpp:3-12; candidates tested in ROI-local coordinates, offset added afterwards, :361-364),
 * then one small download.  pmv_tracker_init is initialise()'s extraction on the first frame (:440-459).
 * Defaults of the reference: win 32x32, max_level 4, min_tracked 400, tracked_tol 150, grid 255, quality 0.01,
 * min_dist 5, neighbor_dist 5. */
typedef struct pmv_tracker pmv_tracker;
PMV_API pmv_tracker *pmv_tracker_create(pmv_ctx *ctx, int rows, int cols, int win_w, int win_h, int max_level, int capacity,
                                        int min_tracked, int tracked_tol, int grid, double quality, double min_dist,
                                        int neighbor_dist);
PMV_API void pmv_tracker_destroy(pmv_tracker *t);
PMV_API int pmv_tracker_init(pmv_tracker *t, const uint8_t *frame, int step, int *n_features);
/* xy: capacity x (column, row) int32 of the new frame's features (tracked ones first, in the order of the previous
 * list, then newly extracted ones); prev_index: position of each in the previous frame's list, -1 = new. */
PMV_API int pmv_tracker_add_frame(pmv_tracker *t, const uint8_t *frame, int step, int *n_tracked, int *n_features,
                                  int *extracted, int32_t *xy, int32_t *prev_index, int capacity);
PMV_API int pmv_tracker_features(pmv_tracker *t, int32_t *xy, int capacity, int *n);

/* ------------------------------------------------------------------ pose from 3-D / 2-D correspondences -- */
/* == cv::solvePnPRansac(obj, img, K, noArray(), rvec, tvec, use_extrinsic_guess, iterations, reproj_err, confidence,
 *    inliers) with the default SOLVEPNP_ITERATIVE flag, as OpenCVEPnPSolver::solvePnP calls it
 * (OpenCVEPnPSolver.cpp:34-35: true, 100, 8, .99): 5-point EPnP hypotheses on cv::RNG((uint64)-1) subsets, float
 * reprojection-error inlier test, RANSACUpdateNumIters, then the minimisation of the reprojection error over the
 * inliers starting from the caller's pose.  obj_xyz: n x 3 float (cv::Point3f), img_xy: n x 2 float; rvec / tvec:
 * in = guess, out = pose (unchanged when no model is found: *n_inliers = 0, OpenCV returns false); inlier_mask:
 * n bytes (optional).  n >= 6. */
PMV_API int pmv_pnp_ransac(pmv_ctx *ctx, const float *obj_xyz, const float *img_xy, int n, const double K[9], double rvec[3],
                           double tvec[3], int use_extrinsic_guess, int iterations, float reproj_err, double confidence,
                           uint8_t *inlier_mask, int *n_inliers);

/* ------------------------------------------------------------------ two-view initialisation -- */
/* The two OpenCV calls of OpenCVFivePointTri::triangulate (OpenCVFivePointTri.cpp:25-27), point lists as n x 2 doubles
 * (pixel coordinates; cv::Point / Point2f / Point2d all convert exactly), K row major, E / R row major.
 *
 * pmv_find_essential_mat == E = cv::findEssentialMat(p1, p2, K, cv::RANSAC, prob, threshold, maxIters, mask)
 * (the reference passes 0.99, 1 and the default 1000): Nister five-point models on cv::RNG((uint64)-1) subsets, Sampson
 * distance against threshold / ((fx + fy) / 2), RANSACUpdateNumIters.  *n_inliers = 0 and E = 0 when no model has
 * more than four inliers (OpenCV returns an empty matrix).  mask: n bytes (optional).  n >= 6. */
PMV_API int pmv_find_essential_mat(pmv_ctx *ctx, const double *p1_xy, const double *p2_xy, int n, const double K[9], double prob,
                                   double threshold, int max_iters, double E[9], uint8_t *mask, int *n_inliers);
/* == cv::recoverPose(E, p1, p2, K, R, t, distance_thresh, mask, tri): decomposeEssentialMat, linear triangulation under
 * the four pose candidates, cheirality vote.  mask: n bytes, in = points to consider (NULL: all), out = points in
 * front of both cameras for the winner; tri: 4 x n homogeneous points (optional); *n_good = votes of the winner. */
PMV_API int pmv_recover_pose(pmv_ctx *ctx, const double E[9], const double *p1_xy, const double *p2_xy, int n, const double K[9],
                             double distance_thresh, double R[9], double t[3], uint8_t *mask, double *tri, int *n_good);
/* Both calls in ONE launch (what OpenCVFivePointTri::triangulate does back to back): ransac_mask = findEssentialMat's
 * mask (optional), mask = recoverPose's in/out mask.  When no model is found *n_inliers = 0 and R / t / tri are left
 * untouched (OpenCV would throw in recoverPose). */
PMV_API int pmv_five_point_pose(pmv_ctx *ctx, const double *p1_xy, const double *p2_xy, int n, const double K[9], double prob,
                                double threshold, int max_iters, double distance_thresh, double E[9], double R[9], double t[3],
                                uint8_t *ransac_mask, uint8_t *mask, double *tri, int *n_inliers, int *n_good);

/* ------------------------------------------------------------------ KITTI wire formats (host only) -- */
/* OdometryPipeline::parsePoses (OdometryPipeline.cpp:525-593): one 3x4 row-major [R|t] per line, at most `stop` lines.
 * R: capacity x 9, t: capacity x 3 (either may be NULL); *n = poses in the file (may exceed capacity).
 * PMV_ERR_INVALID when the file cannot be opened (the reference throws "Unable to open pose file"). */
PMV_API int pmv_kitti_parse_poses(const char *path, int stop, double *R, double *t, int capacity, int *n);
/* OdometryPipeline::parseCalibration (:595-653): the left 3x3 of line `num_calib` ("Pn: ...") of calib.txt into K
 * (row major; entries the line does not reach keep their value, as in the reference). */
PMV_API int pmv_kitti_parse_calibration(const char *path, int num_calib, double K[9]);
/* The error report of OdometryPipeline::run (:272-300): per-frame Frobenius / L2 distances between the n estimated
 * poses (R: n x 9, t: n x 3; index 0 = initial pose, skipped) and the ground truth, with the reference's sign flips
 * and its gt_R[i] (not i + init_offset) quirk.  stats = {R total, min, max, std, t total, min, max, std}; text = the
 * error_path file content (optional). */
PMV_API int pmv_kitti_error_report(const double *R, const double *t, int n, const double *gt_R, const double *gt_t, int n_gt,
                                   int init_offset, double runtime, double stats[8], char *text, int text_capacity);

/* ------------------------------------------------------------------ corner detectors -- */
/* Image arguments of the extractors: `base` is the PARENT image (full_rows x full_cols, row step
 * `step` bytes) and (roi_x, roi_y, roi_w, roi_h) the view the pipeline passes
 * (Frame::regionOfInterest, Frame.cpp:95-117; OdometryPipeline.cpp:674-692 makes 255x255 tiles).
 * Like cv::Sobel on a cv::Mat sub-view, derivative taps read parent pixels beyond the ROI edge
 * and reflect (BORDER_REFLECT_101) only at the parent's edges.  Coordinates returned are
 * ROI-local, as the reference extractors return them. */

/* cv::cornerMinEigenVal(view, eig, 3, 3) -- response stage of goodFeaturesToTrack
 * (OpenCVGoodFeatureExtractor.cpp:7).  eig: roi_h x roi_w floats.  Stage-by-stage parity. */
PMV_API int pmv_min_eigen_val(pmv_ctx *ctx, const uint8_t *base, int full_rows, int full_cols, int step,
                              int roi_x, int roi_y, int roi_w, int roi_h, float *eig);

/* Response stages on `batch` images RESIDENT in HBM (device pointers; asynchronous on the context stream): the
 * cornerMinEigenVal map (OpenCVGoodFeatureExtractor.cpp:7) and the reference's own fp64 response
 * (ShiTomasiFeatureExtractor.cpp:49-75 + Frame.cpp:58-138) of every image, plus each map's maximum (the threshold base
 * of both extractors).  Image b starts at d_imgs + b*img_stride; map b at d_eig / d_R + b*rows*cols. */
PMV_API int pmv_min_eigen_val_batched_dev(pmv_ctx *ctx, const uint8_t *d_imgs, int batch, size_t img_stride, int rows, int cols,
                                          int step, float *d_eig, float *d_max);
PMV_API int pmv_shitomasi_response_batched_dev(pmv_ctx *ctx, const uint8_t *d_imgs, int batch, size_t img_stride, int rows,
                                               int cols, int step, int signed_quirk, double *d_R, double *d_max);

/* == cv::goodFeaturesToTrack(view, corners, max_corners, quality, min_dist, Mat(), block_size,
 *    ksize, false, 0.04) as called at OpenCVGoodFeatureExtractor.cpp:7 (max, 0.01, 5, 3, 3).
 * xy: max_corners x (x,y) floats (integer valued), strongest first; score: the response
 * (the reference adapter leaves Feature::score 0); *n_out corners written.
 * max_corners <= 0 means "all" (size xy for roi_w*roi_h/4).  Only block_size 3 / ksize 3. */
PMV_API int pmv_gftt(pmv_ctx *ctx, const uint8_t *base, int full_rows, int full_cols, int step,
                     int roi_x, int roi_y, int roi_w, int roi_h, int max_corners, double quality,
                     double min_dist, int block_size, int ksize, float *xy, float *score, int *n_out);

/* pmv_gftt on an image that is already RESIDENT in HBM (device pointer, read in place; the fast response kernels need
 * base, step and roi_x 4-byte aligned, otherwise the tile kernels run).  The corner list stays on the device:
 * d_xy (max_corners x 2 floats), d_score (max_corners floats, may be NULL); *n_out on the host -- the call
 * synchronises on the two counts goodFeaturesToTrack's control flow needs (candidates, corners), nothing else crosses
 * PCIe.  max_corners > 0, block_size 3 / ksize 3. */
PMV_API int pmv_gftt_dev(pmv_ctx *ctx, const uint8_t *d_base, int full_rows, int full_cols, int step,
                         int roi_x, int roi_y, int roi_w, int roi_h, int max_corners, double quality,
                         double min_dist, float *d_xy, float *d_score, int *n_out);

/* ShiTomasiFeatureExtractor::computeShiTomasiResponse (ShiTomasiFeatureExtractor.cpp:49-75) on
 * Frame::getHarrisMatrix() (Frame.cpp:58-86, 119-138): fp64 response map, rows x cols doubles.
 * The view is isolated exactly as in the reference (fresh gradient / harris Mats).
 * signed_quirk = 1 reproduces the reference's u8-read-as-schar gradient (Frame.cpp:65-67). */
PMV_API int pmv_shitomasi_response(pmv_ctx *ctx, const uint8_t *img, int rows, int cols, int step,
                                   int signed_quirk, double *R);

/* == ShiTomasiFeatureExtractor::extractFeatures(frame, max) (ShiTomasiFeatureExtractor.cpp:5-47):
 * pixels with R > quality * max(R) (quality 0.4, .h:10), sorted by score descending, first `max`.
 * col/row/score: max entries.  No NMS, like the reference. */
PMV_API int pmv_shitomasi(pmv_ctx *ctx, const uint8_t *img, int rows, int cols, int step, int max_feats,
                          double quality, int signed_quirk, int *col, int *row, double *score, int *n_out);

/* == cv::FAST(view, kp, threshold, nonmax) TYPE_9_16 + "first max in raster order"
 * (OpenCVFASTFeatureExtractor.cpp:8-20; threshold 10, nonmax true).  score = KeyPoint::response.
 * *n_total (optional) = number of keypoints cv::FAST would return. */
PMV_API int pmv_fast(pmv_ctx *ctx, const uint8_t *img, int rows, int cols, int step, int threshold, int nonmax,
                     int max_feats, int *col, int *row, float *score, int *n_out, int *n_total);

/* ------------------------------------------------------------------ bundle adjustment -- */
/* Parameterisation of CeresBundleAdjustment::apply (CeresBundleAdjustment.cpp:26-52):
 *   pose block  [a(3), c(3)]  a = Rodrigues(R_i^T), c = -t_i                       (fp64)
 *   point block X(3) world coordinates
 *   observation (column, row) of the Feature, integer valued; cam_idx / pt_idx name its blocks
 *   K row-major 3x3 (fx = K[0], cx = K[2], fy = K[4], cy = K[5], ProjectionResidual.h:51-52)
 *   loss ceres::HuberLoss(huber_delta) (1.0 in the reference; <= 0 disables the loss). */
typedef struct pmv_ba_summary {
    double initial_cost, final_cost; /* 1/2 sum rho(|r|^2) at the first / last accepted iterate */
    int iterations;                  /* LM iterations executed (accepted + rejected + invalid) */
    int successful_steps;
    int termination;                 /* 0 max iterations, 1 function tol, 2 parameter tol, 3 gradient tol,
                                        4 failure (5 consecutive invalid steps), 5 min trust-region radius */
    double final_radius;
} pmv_ba_summary;

typedef struct pmv_ba_problem pmv_ba_problem;

/* ProjectionResidual::operator() + its AutoDiff Jacobians (include/ProjectionResidual.h:38-58,
 * ProjectionResidual.cpp:3-8) for No observations: r (No x 2), J_pose (No x 2 x 6 row major),
 * J_pt (No x 2 x 3) exactly as the CostFunction returns them (loss not applied), and
 * *cost = 1/2 sum rho(|r|^2).  Any output pointer may be NULL. */
PMV_API int pmv_ba_eval(pmv_ctx *ctx, const double *poses, const double *points, const double *obs,
                        const int32_t *cam_idx, const int32_t *pt_idx, int Nc, int Np, int No, const double K[9],
                        double huber_delta, double *r, double *J_pose, double *J_pt, double *cost);

/* == ceres::Solve with SPARSE_SCHUR, max_num_iterations = max_iters, everything else default
 * (CeresBundleAdjustment.cpp:54-61): Levenberg-Marquardt trust region, Jacobi scaling, points eliminated,
 * reduced camera system solved by Cholesky.  poses (Nc x 6) and points (Np x 3) are updated in place,
 * like tr_opt / p3d_opt are (:67-88).  Blocks without observations are left untouched. */
PMV_API int pmv_ba_solve(pmv_ctx *ctx, double *poses, double *points, const double *obs, const int32_t *cam_idx,
                         const int32_t *pt_idx, int Nc, int Np, int No, const double K[9], double huber_delta,
                         int max_iters, pmv_ba_summary *summary);

/* Diagnostics / tests (no GPU needed): the host side of pmv_ba_problem_create -- the caller's observation list brought
 * into device order, i.e. sorted by (window, point, camera, original index).  Outputs (any may be NULL): pt_off
 * (W*Np + 1), cam_off (W*Nc + 1), cam / pt / win (No each), obs_sorted (2 No), cam_obs (No: device indices grouped by
 * (window, camera), ascending), *route = 0 list taken in place (already ordered), 1 sorted window by window, 2 general. */
PMV_API int pmv_ba_index_observations(const double *obs, const int32_t *cam_idx, const int32_t *pt_idx, const int32_t *obs_off, int W,
                                      int Nc, int Np, int No, int32_t *pt_off, int32_t *cam_off, int32_t *cam, int32_t *pt, int32_t *win,
                                      double *obs_sorted, int32_t *cam_obs, int *route);

/* W independent windows in one call (BASELINE config 4): window w owns poses[w*Nc..], points[w*Np..]
 * and the observation slice [obs_off[w], obs_off[w+1]) whose cam_idx / pt_idx are window-local (obs_off[0] == 0,
 * obs_off[W] == No). */
PMV_API int pmv_ba_solve_batched(pmv_ctx *ctx, double *poses, double *points, const double *obs,
                                 const int32_t *cam_idx, const int32_t *pt_idx, const int32_t *obs_off, int W,
                                 int Nc, int Np, int No, const double K[9], double huber_delta, int max_iters,
                                 pmv_ba_summary *sums);

/* Resident form of the same solve: create uploads and indexes the problem once, solve enqueues
 * max_iters LM iterations on the context stream WITHOUT host synchronisation, download fetches the
 * result (synchronises), reset restores the initial parameters (NULL = the ones given at creation).
 * sharded_nranks > 1: this rank holds all Nc poses but only its shard of the points / observations;
 * every LM iteration sums the partial reduced camera system over ranks (pmv_comm_init first). */
PMV_API pmv_ba_problem *pmv_ba_problem_create(pmv_ctx *ctx, const double *poses, const double *points,
                                              const double *obs, const int32_t *cam_idx, const int32_t *pt_idx,
                                              const int32_t *obs_off, int W, int Nc, int Np, int No,
                                              const double K[9], double huber_delta, int sharded_rank,
                                              int sharded_nranks);
PMV_API int pmv_ba_problem_reset(pmv_ba_problem *p, const double *poses, const double *points);
PMV_API int pmv_ba_problem_solve(pmv_ba_problem *p, int max_iters);
PMV_API int pmv_ba_problem_download(pmv_ba_problem *p, double *poses, double *points, pmv_ba_summary *sums);
PMV_API size_t pmv_ba_problem_device_bytes(pmv_ba_problem *p);
PMV_API void pmv_ba_problem_destroy(pmv_ba_problem *p);

/* NCCL communicator of the sharded bundle adjuster (one rank per GPU): rank 0 makes the id, the
 * caller broadcasts its 128 bytes (e.g. torch.distributed), every rank calls pmv_comm_init. */
PMV_API int pmv_comm_unique_id(char id[128]);
PMV_API int pmv_comm_init(pmv_ctx *ctx, int nranks, int rank, const char id[128]);
PMV_API int pmv_comm_destroy(pmv_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* PMV_CUDA_H */
