// oracle/ref_shim/shim_impl.cpp -- TEST INFRASTRUCTURE ONLY.
// Out-of-line part of the OpenCV / Ceres shim (see opencv2/core.hpp, ceres/ceres.h in this directory): the cv::
// functions that forward to the cv2 hooks (ref_hooks.h) or the plain-C oracle, the small exact ones
// (threshold, cvtColor, Rodrigues), and ceres::Solve.
#include <cfloat>
#include <map>
#include <sstream>

#include <ceres/ceres.h>
#include <opencv2/calib3d.hpp>
#include <opencv2/core.hpp>
#include <opencv2/features2d.hpp>
#include <opencv2/imgproc.hpp>
#include <opencv2/video/tracking.hpp>

#include "ref_hooks.h"

pmv_ref_hooks g_pmv_ref_hooks = {0, 0, 0, 0, 0};

extern "C" {
// plain-C oracle (oracle/pmv_oracle_*.c), linked into the same shared object
int orc_lk_track(const uint8_t* prev, const uint8_t* next, int rows, int cols, int step, const float* pts, int n,
                 int win_w, int win_h, int max_level, int max_count, double eps, int flags, double min_eig,
                 float* out, uint8_t* status, float* err);
int orc_gftt(const uint8_t* img, int full_rows, int full_cols, int step, int rx, int ry, int rw, int rh,
             int max_corners, double quality, double min_dist, float* xy, float* score, int cap);
int orc_fast(const uint8_t* img, int rows, int cols, int step, int threshold, int nonmax, int* col, int* row,
             float* score, int cap);
typedef struct {
    double initial_cost, final_cost; int iterations, successful_steps, termination; double final_radius;
    double cost_log[128], radius_log[128]; int accepted_log[128];
} orc_ba_summary;
typedef void (*orc_residual_hook_t)(const double pose[6], const double pt[3], const double obs[2], const double K[9],
                                    double r[2], double* Jc, double* Jp);
void orc_ba_set_residual_hook(orc_residual_hook_t h);
int orc_ba_solve(double* poses, double* points, const double* obs, const int* cam_idx, const int* pt_idx, int Nc, int Np,
                 int No, const double* K, double huber_delta, int max_iters, int use_dense, orc_ba_summary* out);
}

namespace cv {

static inline int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

// cv::blur(src, dst, Size(3,3)): normalised box filter, BORDER_REFLECT_101, on an isolated Mat.
void blur(const Mat& src, OutputArray dst, Size ksize)
{
    if (src.depth() != CV_64F || ksize.width != 3 || ksize.height != 3) throw std::runtime_error("shim blur: CV_64F 3x3 only");
    const int cn = src.channels(), rows = src.rows, cols = src.cols;
    Mat in = src.clone();                                  // dst may alias src (Frame.cpp:136); contiguous copy for the hook
    Mat out(rows, cols, src.type());
    if (g_pmv_ref_hooks.blur3) {
        if (g_pmv_ref_hooks.blur3(in.ptr<double>(0), rows, cols, cn, out.ptr<double>(0)) != 0) throw std::runtime_error("blur hook failed");
    } else {
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < cols; c++)
                for (int k = 0; k < cn; k++) {
                    double s = 0;
                    for (int dy = -1; dy <= 1; dy++) { const double* p = in.ptr<double>(reflect101(r + dy, rows));
                        for (int dx = -1; dx <= 1; dx++) s += p[reflect101(c + dx, cols) * cn + k]; }
                    out.ptr<double>(r)[c * cn + k] = s * (1.0 / 9.0);
                }
    }
    *dst.m = out;
}

void threshold(const Mat& src, OutputArray dst, double thresh, double maxval, int type)
{
    if (src.type() != CV_64FC1 || type != THRESH_BINARY) throw std::runtime_error("shim threshold: CV_64FC1 THRESH_BINARY only");
    Mat out(src.rows, src.cols, src.type());
    for (int r = 0; r < src.rows; r++) { const double* s = src.ptr<double>(r); double* d = out.ptr<double>(r);
        for (int c = 0; c < src.cols; c++) d[c] = s[c] > thresh ? maxval : 0.0; }      // NaN > thresh is false, as in OpenCV
    *dst.m = out;
}

void cvtColor(const Mat& src, OutputArray dst, int code)
{
    if (code != COLOR_BGR2GRAY || src.type() != CV_8UC3) throw std::runtime_error("shim cvtColor: 8-bit BGR2GRAY only");
    Mat out(src.rows, src.cols, CV_8UC1);
    for (int r = 0; r < src.rows; r++) { const uchar* s = src.ptr<uchar>(r); uchar* d = out.ptr<uchar>(r);
        // OpenCV's 8-bit path: fixed point, 14 fractional bits, B 0.114 / G 0.587 / R 0.299
        for (int c = 0; c < src.cols; c++) d[c] = (uchar)((s[3 * c] * 1868 + s[3 * c + 1] * 9617 + s[3 * c + 2] * 4899 + (1 << 13)) >> 14); }
    *dst.m = out;
}

void goodFeaturesToTrack(const Mat& image, std::vector<Point2f>& corners, int maxCorners, double qualityLevel, double minDistance,
                         const Mat& mask, int blockSize, int gradientSize, bool useHarris, double)
{
    if (image.type() != CV_8UC1 || !mask.empty() || blockSize != 3 || gradientSize != 3 || useHarris)
        throw std::runtime_error("shim goodFeaturesToTrack: the reference's arguments only");
    Size whole; Point ofs; image.locateROI(whole, ofs);
    const uchar* base = image.data - (size_t)ofs.y * image.step - ofs.x;
    const int cap = maxCorners > 0 ? maxCorners : image.rows * image.cols;
    std::vector<float> xy(2 * (size_t)cap + 2), score(cap + 1);
    int n = 0;
    if (g_pmv_ref_hooks.gftt) {
        if (g_pmv_ref_hooks.gftt(base, whole.height, whole.width, (int)image.step, ofs.x, ofs.y, image.cols, image.rows, maxCorners,
                                 qualityLevel, minDistance, cap, xy.data(), &n) != 0) throw std::runtime_error("gftt hook failed");
    } else
        n = orc_gftt(base, whole.height, whole.width, (int)image.step, ofs.x, ofs.y, image.cols, image.rows, maxCorners,
                     qualityLevel, minDistance, xy.data(), score.data(), cap);
    corners.clear();
    for (int i = 0; i < n; i++) corners.push_back(Point2f(xy[2 * i], xy[2 * i + 1]));
}

void FAST(const Mat& image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmax)
{
    if (image.type() != CV_8UC1) throw std::runtime_error("shim FAST: CV_8UC1 only");
    const int cap = image.rows * image.cols;
    keypoints.clear();
    int n = 0;
    if (g_pmv_ref_hooks.fast) {
        std::vector<float> xy(2 * (size_t)cap), resp(cap);
        if (g_pmv_ref_hooks.fast(image.data, image.rows, image.cols, (int)image.step, threshold, nonmax, cap, xy.data(), resp.data(), &n) != 0)
            throw std::runtime_error("fast hook failed");
        for (int i = 0; i < n; i++) { KeyPoint k; k.pt = Point2f(xy[2 * i], xy[2 * i + 1]); k.size = 7.f; k.response = resp[i]; keypoints.push_back(k); }
    } else {
        std::vector<int> col(cap), row(cap); std::vector<float> sc(cap);
        n = orc_fast(image.data, image.rows, image.cols, (int)image.step, threshold, nonmax, col.data(), row.data(), sc.data(), cap);
        if (n > cap) n = cap;
        for (int i = 0; i < n; i++) { KeyPoint k; k.pt = Point2f((float)col[i], (float)row[i]); k.size = 7.f; k.response = sc[i]; keypoints.push_back(k); }
    }
}

void calcOpticalFlowPyrLK(const Mat& prevImg, const Mat& nextImg, const std::vector<Point2f>& prevPts, std::vector<Point2f>& nextPts,
                          std::vector<uchar>& status, std::vector<float>& err, Size winSize, int maxLevel, TermCriteria crit, int flags,
                          double minEig)
{
    if (prevImg.type() != CV_8UC1 || nextImg.type() != CV_8UC1 || prevImg.rows != nextImg.rows || prevImg.cols != nextImg.cols || flags != 0)
        throw std::runtime_error("shim calcOpticalFlowPyrLK: equal-size CV_8UC1, flags 0 only");
    const int n = (int)prevPts.size();
    nextPts.assign(n, Point2f()); status.assign(n, 0); err.assign(n, 0.f);
    if (!n) return;
    std::vector<float> in(2 * (size_t)n), out(2 * (size_t)n);
    for (int i = 0; i < n; i++) { in[2 * i] = prevPts[i].x; in[2 * i + 1] = prevPts[i].y; }
    if (g_pmv_ref_hooks.lk) {
        if (g_pmv_ref_hooks.lk(prevImg.data, nextImg.data, prevImg.rows, prevImg.cols, (int)prevImg.step, (int)nextImg.step, in.data(), n,
                               winSize.width, winSize.height, maxLevel, out.data(), status.data(), err.data()) != 0)
            throw std::runtime_error("lk hook failed");
    } else {
        Mat a = prevImg, b = nextImg;
        if (a.step != b.step) { a = a.clone(); b = b.clone(); }
        orc_lk_track(a.data, b.data, a.rows, a.cols, (int)a.step, in.data(), n, winSize.width, winSize.height, maxLevel, crit.maxCount,
                     crit.epsilon, flags, minEig, out.data(), status.data(), err.data());
    }
    for (int i = 0; i < n; i++) nextPts[i] = Point2f(out[2 * i], out[2 * i + 1]);
}

// cv::Rodrigues for proper rotations: matrix -> vector by the logarithm map, vector -> matrix by Rodrigues' formula
// (OpenCV additionally re-orthonormalises the input matrix by SVD; for the exactly-orthonormal inputs used here the
// two agree to rounding.  Both the reference class and its GPU replacement run through this same function.)
void Rodrigues(const Mat& src, OutputArray dst)
{
    if (src.depth() != CV_64F) throw std::runtime_error("shim Rodrigues: CV_64F only");
    if (src.rows == 3 && src.cols == 3) {
        double R[9]; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R[3 * i + j] = src.at<double>(i, j);
        double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
        double s = std::sqrt((rx * rx + ry * ry + rz * rz) * 0.25), c = (R[0] + R[4] + R[8] - 1) * 0.5;
        c = c > 1. ? 1. : c < -1. ? -1. : c;
        double theta = std::acos(c), v[3];
        if (s < 1e-5) {
            if (c > 0) v[0] = v[1] = v[2] = 0;
            else {
                double t;
                t = (R[0] + 1) * 0.5; v[0] = std::sqrt(std::max(t, 0.));
                t = (R[4] + 1) * 0.5; v[1] = std::sqrt(std::max(t, 0.)) * (R[1] < 0 ? -1. : 1.);
                t = (R[8] + 1) * 0.5; v[2] = std::sqrt(std::max(t, 0.)) * (R[2] < 0 ? -1. : 1.);
                if (std::fabs(v[0]) < std::fabs(v[1]) && std::fabs(v[0]) < std::fabs(v[2]) && (R[5] > 0) != (v[1] * v[2] > 0)) v[2] = -v[2];
                double nrm = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); theta /= nrm;
                for (int i = 0; i < 3; i++) v[i] *= theta;
            }
        } else { double vth = 1 / (2 * s) * theta; v[0] = rx * vth; v[1] = ry * vth; v[2] = rz * vth; }
        Mat out(3, 1, CV_64FC1); for (int i = 0; i < 3; i++) out.at<double>(i) = v[i];
        if (dst.m->data && dst.m->rows * dst.m->cols == 3 && dst.m->type() == CV_64FC1) for (int i = 0; i < 3; i++) dst.m->at<double>(i) = v[i];
        else *dst.m = out;
    } else if (src.rows * src.cols == 3) {
        double r[3] = {src.at<double>(0), src.at<double>(1), src.at<double>(2)}, R[9];
        double theta = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
        if (theta < DBL_EPSILON) { for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0); }
        else {
            double c = std::cos(theta), s = std::sin(theta), c1 = 1. - c, it = 1. / theta;
            double x = r[0] * it, y = r[1] * it, z = r[2] * it;
            double rrt[9] = {x * x, x * y, x * z, x * y, y * y, y * z, x * z, y * z, z * z};
            double rx[9] = {0, -z, y, z, 0, -x, -y, x, 0};
            for (int k = 0; k < 9; k++) R[k] = c * (k % 4 == 0) + c1 * rrt[k] + s * rx[k];
        }
        if (dst.m->data && dst.m->rows == 3 && dst.m->cols == 3 && dst.m->type() == CV_64FC1)
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) dst.m->at<double>(i, j) = R[3 * i + j];
        else { Mat out(3, 3, CV_64FC1); for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out.at<double>(i, j) = R[3 * i + j]; *dst.m = out; }
    } else throw std::runtime_error("shim Rodrigues: 3x3 or 3-vector only");
}

bool solvePnPRansac(const std::vector<Point3f>& obj, const std::vector<Point2f>& img, const Mat& K, const Mat& dist, OutputArray rvec,
                    OutputArray tvec, bool useGuess, int iters, float reprojErr, double confidence, std::vector<int>& inliers)
{
    if (!g_pmv_ref_hooks.pnp_ransac) throw std::runtime_error("shim solvePnPRansac: needs the cv2 hook");
    if (!dist.empty()) throw std::runtime_error("shim solvePnPRansac: no distortion");
    const int n = (int)obj.size();
    std::vector<float> o(3 * (size_t)n), p(2 * (size_t)n);
    for (int i = 0; i < n; i++) { o[3 * i] = obj[i].x; o[3 * i + 1] = obj[i].y; o[3 * i + 2] = obj[i].z; p[2 * i] = img[i].x; p[2 * i + 1] = img[i].y; }
    double Kd[9], rv[3] = {0, 0, 0}, tv[3] = {0, 0, 0};
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Kd[3 * i + j] = K.at<double>(i, j);
    if (useGuess) for (int i = 0; i < 3; i++) { rv[i] = rvec.m->at<double>(i); tv[i] = tvec.m->at<double>(i); }
    std::vector<int> inl(n + 1); int ninl = 0;
    int ok = g_pmv_ref_hooks.pnp_ransac(o.data(), p.data(), n, Kd, rv, tv, useGuess, iters, reprojErr, confidence, inl.data(), &ninl);
    Mat r(3, 1, CV_64FC1), t(3, 1, CV_64FC1);
    for (int i = 0; i < 3; i++) { r.at<double>(i) = rv[i]; t.at<double>(i) = tv[i]; }
    *rvec.m = r; *tvec.m = t;
    inliers.assign(inl.begin(), inl.begin() + ninl);
    return ok != 0;
}

static void shim_points(const std::vector<Point>& p, std::vector<double>& out)
{
    out.resize(2 * p.size());
    for (size_t i = 0; i < p.size(); i++) { out[2 * i] = p[i].x; out[2 * i + 1] = p[i].y; }
}

Mat findEssentialMat(const std::vector<Point>& p1, const std::vector<Point>& p2, const Mat& K, int method, double prob, double threshold,
                     OutputArray mask)
{
    if (!g_pmv_ref_hooks.find_essential) throw std::runtime_error("shim findEssentialMat: needs the cv2 hook");
    const int n = (int)p1.size();
    std::vector<double> a, b; shim_points(p1, a); shim_points(p2, b);
    double Kd[9], E[90];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Kd[3 * i + j] = K.at<double>(i, j);
    Mat m(n, 1, CV_8UC1);
    const int rows = g_pmv_ref_hooks.find_essential(a.data(), b.data(), n, Kd, method, prob, threshold, E, m.data);
    if (rows < 0) throw std::runtime_error("shim findEssentialMat: cv2 failed");
    *mask.m = m;
    if (rows == 0) return Mat();
    Mat Em(rows, 3, CV_64FC1);
    for (int i = 0; i < rows; i++) for (int j = 0; j < 3; j++) Em.at<double>(i, j) = E[3 * i + j];
    return Em;
}

int recoverPose(const Mat& E, const std::vector<Point>& p1, const std::vector<Point>& p2, const Mat& K, OutputArray R, OutputArray t,
                double distanceThresh, OutputArray mask, OutputArray triangulatedPoints)
{
    if (!g_pmv_ref_hooks.recover_pose) throw std::runtime_error("shim recoverPose: needs the cv2 hook");
    if (E.rows != 3 || E.cols != 3) throw std::runtime_error("shim recoverPose: E must be 3x3");
    const int n = (int)p1.size();
    std::vector<double> a, b; shim_points(p1, a); shim_points(p2, b);
    double Kd[9], Ed[9], Rd[9], td[3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { Kd[3 * i + j] = K.at<double>(i, j); Ed[3 * i + j] = E.at<double>(i, j); }
    Mat tri(4, n, CV_64FC1);
    if (mask.m->empty()) { Mat m(n, 1, CV_8UC1); std::memset(m.data, 1, n); *mask.m = m; }
    const int good = g_pmv_ref_hooks.recover_pose(Ed, a.data(), b.data(), n, Kd, distanceThresh, Rd, td, mask.m->data, tri.ptr<double>(0));
    if (good < 0) throw std::runtime_error("shim recoverPose: cv2 failed");
    Mat Rm(3, 3, CV_64FC1), tm(3, 1, CV_64FC1);
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) Rm.at<double>(i, j) = Rd[3 * i + j]; tm.at<double>(i) = td[i]; }
    *R.m = Rm; *t.m = tm; *triangulatedPoints.m = tri;
    return good;
}
}  // namespace cv

// ---------------------------------------------------------------------------------------------- ceres::Solve
namespace ceres {
static Solver::Summary g_last;
const Solver::Summary& LastSummary() { return g_last; }
std::string Solver::Summary::FullReport() const
{
    std::ostringstream s; s << "oracle LM (Ceres shim): cost " << initial_cost << " -> " << final_cost << " in " << iterations << " iterations";
    return s.str();
}

static const std::vector<Problem::ResidualBlock>* g_blocks = nullptr;
static void eval_hook(const double pose[6], const double pt[3], const double obs[2], const double*, double r[2], double* Jc, double* Jp)
{   // obs[0] carries the residual-block id (see orc_ba_set_residual_hook)
    const Problem::ResidualBlock& b = (*g_blocks)[(size_t)obs[0]];
    const double* params[2] = {pose, pt};
    double* jac[2] = {Jc, Jp};
    b.cost->Evaluate(params, r, (Jc || Jp) ? jac : nullptr);
}

// Options honoured: max_num_iterations.  linear_solver_type must be SPARSE_SCHUR (what the reference sets); all other
// minimiser settings are Ceres' defaults, restated in oracle/pmv_oracle_ba.c.
void Solve(const Solver::Options& options, Problem* problem, Solver::Summary* summary)
{
    if (options.linear_solver_type != SPARSE_SCHUR) throw std::runtime_error("ceres shim: SPARSE_SCHUR only");
    const auto& blocks = problem->blocks();
    std::map<double*, int> cam_slot, pt_slot;               // parameter blocks by address, in order of first use
    std::vector<double*> cam_ptr, pt_ptr;
    std::vector<int> cam_idx, pt_idx; std::vector<double> payload;
    double delta = -1.0;
    for (size_t i = 0; i < blocks.size(); i++) {
        const auto& b = blocks[i];
        if (b.cost->parameter_block_sizes() != std::vector<int>({6, 3}) || b.cost->num_residuals() != 2)
            throw std::runtime_error("ceres shim: 2 residuals over a 6-block and a 3-block only");
        if (!cam_slot.count(b.x0)) { cam_slot[b.x0] = (int)cam_ptr.size(); cam_ptr.push_back(b.x0); }
        if (!pt_slot.count(b.x1)) { pt_slot[b.x1] = (int)pt_ptr.size(); pt_ptr.push_back(b.x1); }
        cam_idx.push_back(cam_slot[b.x0]); pt_idx.push_back(pt_slot[b.x1]);
        payload.push_back((double)i); payload.push_back(0.0);
        const double d = b.loss ? b.loss->huber_delta() : -1.0;
        if (i && d != delta) throw std::runtime_error("ceres shim: one loss for all blocks");
        delta = d;
    }
    const int Nc = (int)cam_ptr.size(), Np = (int)pt_ptr.size(), No = (int)blocks.size();
    std::vector<double> poses(6 * (size_t)Nc + 1), points(3 * (size_t)Np + 1);
    for (int c = 0; c < Nc; c++) std::memcpy(&poses[6 * c], cam_ptr[c], 6 * sizeof(double));
    for (int p = 0; p < Np; p++) std::memcpy(&points[3 * p], pt_ptr[p], 3 * sizeof(double));
    const double K[9] = {0};
    orc_ba_summary s; std::memset(&s, 0, sizeof s);
    g_blocks = &blocks;
    orc_ba_set_residual_hook(eval_hook);
    orc_ba_solve(poses.data(), points.data(), payload.data(), cam_idx.data(), pt_idx.data(), Nc, Np, No, K, delta,
                 options.max_num_iterations, 0, &s);
    orc_ba_set_residual_hook(nullptr);
    g_blocks = nullptr;
    for (int c = 0; c < Nc; c++) std::memcpy(cam_ptr[c], &poses[6 * c], 6 * sizeof(double));
    for (int p = 0; p < Np; p++) std::memcpy(pt_ptr[p], &points[3 * p], 3 * sizeof(double));
    Solver::Summary out; out.initial_cost = s.initial_cost; out.final_cost = s.final_cost; out.iterations = s.iterations;
    out.num_successful_steps = s.successful_steps; out.termination = s.termination; out.final_radius = s.final_radius;
    g_last = out;
    if (summary) *summary = out;
}
}  // namespace ceres
