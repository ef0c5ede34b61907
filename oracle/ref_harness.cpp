// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// C entry points of oracle/_ref/libpmv_ref.so.  The shared object contains the REFERENCE's own translation units,
// compiled unchanged from /root/reference by oracle/ref_build.py:
//     Feature.cpp Feature3D.cpp Frame.cpp ShiTomasiFeatureExtractor.cpp ProjectionResidual.cpp
//     CeresBundleAdjustment.cpp OpenCVGoodFeatureExtractor.cpp OpenCVFASTFeatureExtractor.cpp OpenCVLucasKanadeFM.cpp
//     OpenCVEPnPSolver.cpp OpenCVFivePointTri.cpp
// against the functional OpenCV / Ceres / dlib shim in oracle/ref_shim/ (third-party kernels forwarded to the real cv2
// wheel through hooks), plus -- in the same object, against the same shim -- the product's drop-in adapters
// (practical-multi-view_b200/host/pmv_adapters.h -> libpmv_cuda.so).  Every entry point takes `impl`:
//     0 = the reference class, 1 = the Gpu* adapter,
// builds the same Frame / OdometryPipeline state for either and returns what the plugin call produced, so the tests
// compare the reference's code with the product through the reference's own plugin interfaces.
#include <cstring>
#include <string>

#include "CeresBundleAdjustment.h"
#include "OdometryPipeline.h"
#include "OpenCVEPnPSolver.h"
#include "OpenCVFASTFeatureExtractor.h"
#include "OpenCVFivePointTri.h"
#include "OpenCVGoodFeatureExtractor.h"
#include "OpenCVLucasKanadeFM.h"
#include "ProjectionResidual.h"
#include "ShiTomasiFeatureExtractor.h"
#include "pmv_adapters.h"
#include "ref_hooks.h"

#define REF_API extern "C" __attribute__((visibility("default")))

static std::string g_err;
#define REF_TRY try {
#define REF_CATCH } catch (const std::exception& e) { g_err = e.what(); return -1; } catch (...) { g_err = "unknown exception"; return -1; }

REF_API const char* ref_last_error() { return g_err.c_str(); }
REF_API void ref_set_hooks(const pmv_ref_hooks* h) { if (h) g_pmv_ref_hooks = *h; else std::memset(&g_pmv_ref_hooks, 0, sizeof g_pmv_ref_hooks); }
REF_API const char* ref_sources()
{
    return "Feature.cpp Feature3D.cpp Frame.cpp ShiTomasiFeatureExtractor.cpp ProjectionResidual.cpp CeresBundleAdjustment.cpp "
           "OpenCVGoodFeatureExtractor.cpp OpenCVFASTFeatureExtractor.cpp OpenCVLucasKanadeFM.cpp OpenCVEPnPSolver.cpp OpenCVFivePointTri.cpp";
}

// Frame(cv::Mat& orig) runs cvtColor(BGR2GRAY) (Frame.cpp:38-42): feed it B = G = R = gray, which the 8-bit fixed-point
// conversion maps back to gray exactly (1868 + 9617 + 4899 = 2^14).
static cv::Mat bgr_of(const uint8_t* gray, int rows, int cols, int step)
{
    cv::Mat m(rows, cols, CV_8UC3);
    for (int r = 0; r < rows; r++) { uint8_t* d = m.ptr<uint8_t>(r); const uint8_t* s = gray + (size_t)r * step;
        for (int c = 0; c < cols; c++) d[3 * c] = d[3 * c + 1] = d[3 * c + 2] = s[c]; }
    return m;
}

// ---- a6-a8: Frame::computeSpatialGradient / computeHarrisMatrix + ShiTomasiFeatureExtractor::computeShiTomasiResponse
REF_API int ref_shitomasi_response(int impl, const uint8_t* gray, int rows, int cols, int step, double* R)
{
    REF_TRY
    cv::Mat bgr = bgr_of(gray, rows, cols, step);
    Frame f(bgr);
    cv::Mat out;
    if (impl == 0) { ShiTomasiFeatureExtractor e; out = e.computeShiTomasiResponse(f); }
    else { GpuShiTomasiFeatureExtractor e; out = e.computeShiTomasiResponse(f); }
    for (int r = 0; r < rows; r++) std::memcpy(R + (size_t)r * cols, out.ptr<double>(r), sizeof(double) * cols);
    return 0;
    REF_CATCH
}

// intermediate planes of the reference Frame (a7 / a8): gradient x, y (rows x cols) and the blurred 3-channel harris
REF_API int ref_frame_planes(const uint8_t* gray, int rows, int cols, int step, double* gx, double* gy, double* harris3)
{
    REF_TRY
    cv::Mat bgr = bgr_of(gray, rows, cols, step);
    Frame f(bgr);
    cv::Mat& X = f.getSpatialGradientX(); cv::Mat& Y = f.getSpatialGradientY(); cv::Mat& Hm = f.getHarrisMatrix();
    for (int r = 0; r < rows; r++) {
        if (gx) std::memcpy(gx + (size_t)r * cols, X.ptr<double>(r), sizeof(double) * cols);
        if (gy) std::memcpy(gy + (size_t)r * cols, Y.ptr<double>(r), sizeof(double) * cols);
        if (harris3) std::memcpy(harris3 + (size_t)r * cols * 3, Hm.ptr<double>(r), sizeof(double) * cols * 3);
    }
    return 0;
    REF_CATCH
}

// ---- a5 / a9 / a10 (+ a15): extractor->extractFeatures(frame-or-ROI, max) as OdometryPipeline.cpp:357 / :450 call it.
// which: 0 ShiTomasiFeatureExtractor, 1 OpenCVGoodFeatureExtractor, 2 OpenCVFASTFeatureExtractor.
// roi = {x, y, w, h} or NULL.  Outputs per feature: column, row, score, tracked, detector.  Returns the count (<= cap).
REF_API int ref_extract(int which, int impl, const uint8_t* gray, int rows, int cols, int step, const int* roi, int max,
                        int* col, int* row, double* score, int* tracked, int cap)
{
    REF_TRY
    cv::Mat bgr = bgr_of(gray, rows, cols, step);
    Frame full(bgr);
    cv::Rect rect = roi ? cv::Rect(roi[0], roi[1], roi[2], roi[3]) : cv::Rect(0, 0, cols, rows);
    Frame view = roi ? full.regionOfInterest(rect) : full;
    std::unique_ptr<BaseFeatureExtractor> e;
    if (impl == 0) {
        if (which == 0) e.reset(new ShiTomasiFeatureExtractor());
        else if (which == 1) e.reset(new OpenCVGoodFeatureExtractor());
        else e.reset(new OpenCVFASTFeatureExtractor());
    } else {
        if (which == 0) e.reset(new GpuShiTomasiFeatureExtractor());
        else if (which == 1) e.reset(new GpuGoodFeatureExtractor());
        else e.reset(new GpuFASTFeatureExtractor());
    }
    std::vector<Feature> feats = e->extractFeatures(view, max);
    int n = 0;
    for (auto& f : feats) {
        if (n >= cap) break;
        col[n] = f.column; row[n] = f.row; score[n] = f.score; tracked[n] = f.tracked ? 1 : 0; n++;
    }
    return n;
    REF_CATCH
}

// ---- a1: matcher->matchFeatures(src, next) as OdometryPipeline.cpp:335 calls it.  src.map is filled with the n features
// (column, row) in the given order; returns the correspondences (src column,row -> next column,row) and next.map's size.
REF_API int ref_match(int impl, const uint8_t* prev, const uint8_t* next, int rows, int cols, int step, const int* feat_cr, int n,
                      int* corr /* cap x 4 */, int cap, int* next_map_size)
{
    REF_TRY
    cv::Mat a = bgr_of(prev, rows, cols, step), b = bgr_of(next, rows, cols, step);
    Frame src(a), dst(b);
    std::vector<std::shared_ptr<Feature3D>> keep;
    for (int i = 0; i < n; i++) src.map[std::make_shared<Feature>(Feature(feat_cr[2 * i], feat_cr[2 * i + 1]))] = std::weak_ptr<Feature3D>();
    std::unique_ptr<BaseFeatureMatcher> m;
    if (impl == 0) m.reset(new OpenCVLucasKanadeFM()); else m.reset(new GpuLucasKanadeFM());
    BaseFeatureMatcher::fmap c = m->matchFeatures(src, dst);
    int k = 0;
    for (auto& p : c) {
        if (k >= cap) break;
        std::shared_ptr<Feature> f0 = p.first.lock(), f1 = p.second.lock();
        if (!f0 || !f1) continue;
        corr[4 * k] = f0->column; corr[4 * k + 1] = f0->row; corr[4 * k + 2] = f1->column; corr[4 * k + 3] = f1->row; k++;
    }
    if (next_map_size) *next_map_size = (int)dst.map.size();
    return k;
    REF_CATCH
}

// ---- a11 / a12: ProjectionResidual::Create(p2d, camera) -> CostFunction::Evaluate (AutoDiff over the reference functor)
REF_API int ref_residual(const double* pose, const double* point, const double* obs, const double* K, int n,
                         const int* cam_idx, const int* pt_idx, double* r, double* Jc, double* Jp)
{
    REF_TRY
    for (int i = 0; i < n; i++) {
        ceres::CostFunction* cf = ProjectionResidual::Create(obs + 2 * i, K);
        const double* params[2] = {pose + 6 * (cam_idx ? cam_idx[i] : i), point + 3 * (pt_idx ? pt_idx[i] : i)};
        double* jac[2] = {Jc ? Jc + 12 * (size_t)i : nullptr, Jp ? Jp + 6 * (size_t)i : nullptr};
        bool ok = cf->Evaluate(params, r + 2 * (size_t)i, (Jc || Jp) ? jac : nullptr);
        delete cf;
        if (!ok) { g_err = "Evaluate returned false"; return -1; }
    }
    return 0;
    REF_CATCH
}

// ---- a13: optimizer->apply(frame) as OdometryPipeline.cpp:410 calls it, on a synthetic pipeline state:
// n_frames poses (R row-major 3x3, t 3), n_points Feature3D (float xyz), observations (frame, point, column, row).
// R / t / points are updated in place, exactly as the plugin leaves OdometryPipeline::R, ::t and the Feature3D objects.
// summary: initial_cost, final_cost, iterations, successful steps.
REF_API int ref_ba_apply(int impl, int n_frames, int bundle_size, int ba_iterations, const double* K, double* R, double* t,
                         int n_points, float* points, int n_obs, const int* obs_frame, const int* obs_point, const int* obs_col,
                         const int* obs_row, int apply_frame, double* summary)
{
    REF_TRY
    OdometryPipeline pipe;
    pipe.bundle_size = bundle_size; pipe.ba_iterations = ba_iterations; pipe.verbose = false;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) pipe.camera.at<double>(i, j) = K[3 * i + j];
    for (int p = 0; p < n_points; p++) pipe.feats3d.push_back(std::make_shared<Feature3D>(cv::Point3f(points[3 * p], points[3 * p + 1], points[3 * p + 2])));
    cv::Mat tiny(1, 1, CV_8UC3);
    for (int i = 0; i < n_frames; i++) {
        auto fr = std::make_shared<Frame>(tiny); fr->frame = i; pipe.frames.push_back(fr);
        cv::Mat Ri(3, 3, CV_64FC1), ti(3, 1, CV_64FC1);
        for (int a = 0; a < 3; a++) { for (int b = 0; b < 3; b++) Ri.at<double>(a, b) = R[9 * i + 3 * a + b]; ti.at<double>(a) = t[3 * i + a]; }
        pipe.R.push_back(Ri); pipe.t.push_back(ti);
    }
    for (int k = 0; k < n_obs; k++)
        pipe.frames[obs_frame[k]]->map[std::make_shared<Feature>(Feature(obs_col[k], obs_row[k]))] = std::weak_ptr<Feature3D>(pipe.feats3d[obs_point[k]]);
    if (impl == 0) {
        CeresBundleAdjustment ba(&pipe); ba.apply(*pipe.frames[apply_frame]);
        const ceres::Solver::Summary& s = ceres::LastSummary();
        if (summary) { summary[0] = s.initial_cost; summary[1] = s.final_cost; summary[2] = s.iterations; summary[3] = s.num_successful_steps; }
    } else {
        GpuBundleAdjustment ba(&pipe); ba.apply(*pipe.frames[apply_frame]);
        if (summary) { summary[0] = ba.last_summary.initial_cost; summary[1] = ba.last_summary.final_cost; summary[2] = ba.last_summary.iterations;
                       summary[3] = ba.last_summary.successful_steps; }
    }
    for (int i = 0; i < n_frames; i++)
        for (int a = 0; a < 3; a++) { for (int b = 0; b < 3; b++) R[9 * i + 3 * a + b] = pipe.R[i].at<double>(a, b); t[3 * i + a] = pipe.t[i].at<double>(a); }
    for (int p = 0; p < n_points; p++) { cv::Point3f q = pipe.feats3d[p]->getPoint(); points[3 * p] = q.x; points[3 * p + 1] = q.y; points[3 * p + 2] = q.z; }
    return 0;
    REF_CATCH
}

// ---- f2: pnpsolver->solvePnP(src, next, R, t) as OdometryPipeline::estimatePose calls it, on a synthetic pipeline state:
// n 3-D points in WORLD coordinates (Feature3D, float), src frame index 1 with pose (R1, t1), each point tracked from
// (src_col, src_row) to (next_col, next_row).  R / t: in = guess, out = estimate; kept[i] = point i survived the outlier
// removal (still in OdometryPipeline::feats3d).
REF_API int ref_pnp_solve(int impl, const double* K, const double* R1, const double* t1, int n, const float* points,
                          const int* src_cr, const int* next_cr, double* R, double* t, int* kept)
{
    REF_TRY
    OdometryPipeline pipe;
    pipe.verbose = false;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) pipe.camera.at<double>(i, j) = K[3 * i + j];
    cv::Mat tiny(1, 1, CV_8UC3);
    Frame src(tiny), next(tiny);
    src.frame = 1; next.frame = 2;
    cv::Mat I(3, 3, CV_64FC1), z(3, 1, CV_64FC1), Rs(3, 3, CV_64FC1), ts(3, 1, CV_64FC1);
    for (int a = 0; a < 3; a++) {
        for (int b = 0; b < 3; b++) { Rs.at<double>(a, b) = R1[3 * a + b]; I.at<double>(a, b) = a == b ? 1.0 : 0.0; }
        ts.at<double>(a) = t1[a]; z.at<double>(a) = 0.0;
    }
    pipe.R.push_back(I); pipe.t.push_back(z); pipe.R.push_back(Rs); pipe.t.push_back(ts);
    std::vector<std::shared_ptr<Feature>> keep_alive;
    for (int p = 0; p < n; p++) {
        auto f3 = std::make_shared<Feature3D>(cv::Point3f(points[3 * p], points[3 * p + 1], points[3 * p + 2]));
        pipe.feats3d.push_back(f3);
        auto fs = std::make_shared<Feature>(Feature(src_cr[2 * p], src_cr[2 * p + 1]));
        auto fn = std::make_shared<Feature>(Feature(next_cr[2 * p], next_cr[2 * p + 1]));
        keep_alive.push_back(fs); keep_alive.push_back(fn);
        src.map[fs] = std::weak_ptr<Feature3D>(f3);
        src.feat_corr[fs] = fn;
    }
    std::vector<std::shared_ptr<Feature3D>> all(pipe.feats3d.begin(), pipe.feats3d.end());
    cv::Mat Rm(3, 3, CV_64FC1), tm(3, 1, CV_64FC1);
    for (int a = 0; a < 3; a++) { for (int b = 0; b < 3; b++) Rm.at<double>(a, b) = R[3 * a + b]; tm.at<double>(a) = t[a]; }
    std::unique_ptr<BasePnPSolver> s;
    if (impl == 0) s.reset(new OpenCVEPnPSolver(&pipe)); else s.reset(new GpuEPnPSolver(&pipe));
    s->solvePnP(src, next, Rm, tm);
    for (int a = 0; a < 3; a++) { for (int b = 0; b < 3; b++) R[3 * a + b] = Rm.at<double>(a, b); t[a] = tm.at<double>(a); }
    for (int p = 0; p < n; p++) kept[p] = std::find(pipe.feats3d.begin(), pipe.feats3d.end(), all[p]) != pipe.feats3d.end();
    return (int)next.map.size();
    REF_CATCH
}

// ---- f4: triangulator->triangulate(src, next, R, t) as OdometryPipeline::initialise calls it, on a synthetic pipeline
// state: n tracked features (src_cr -> next_cr, integer pixels), src frame 0 at the identity pose, ground-truth
// translations gt0 / gt1 (their distance is the scale the reference applies).  Out: R, t (scaled), scale, for every input
// pair the index of its Feature3D in OdometryPipeline::feats3d (-1: none) and the world coordinates of those points
// (points: capacity n x 3).  Returns the number of Feature3D created.
REF_API int ref_triangulate(int impl, const double* K, int n, const int* src_cr, const int* next_cr, const double* gt0, const double* gt1,
                            double* R, double* t, double* scale, int* f3d_index, float* points)
{
    REF_TRY
    OdometryPipeline pipe;
    pipe.verbose = false;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) pipe.camera.at<double>(i, j) = K[3 * i + j];
    cv::Mat tiny(1, 1, CV_8UC3);
    Frame src(tiny), next(tiny);
    src.frame = 0; next.frame = 1;
    cv::Mat I(3, 3, CV_64FC1), z(3, 1, CV_64FC1), g0(3, 1, CV_64FC1), g1(3, 1, CV_64FC1);
    for (int a = 0; a < 3; a++) {
        for (int b = 0; b < 3; b++) I.at<double>(a, b) = a == b ? 1.0 : 0.0;
        z.at<double>(a) = 0.0; g0.at<double>(a) = gt0[a]; g1.at<double>(a) = gt1[a];
    }
    pipe.R.push_back(I); pipe.t.push_back(z);
    pipe.gt_t.push_back(g0); pipe.gt_t.push_back(g1);
    std::vector<std::shared_ptr<Feature>> fs(n), fn(n);
    for (int p = 0; p < n; p++) {
        fs[p] = std::make_shared<Feature>(Feature(src_cr[2 * p], src_cr[2 * p + 1]));
        fn[p] = std::make_shared<Feature>(Feature(next_cr[2 * p], next_cr[2 * p + 1]));
        src.feat_corr[fs[p]] = fn[p];
    }
    cv::Mat Rm, tm;
    std::unique_ptr<BaseTriangulator> tr;
    if (impl == 0) tr.reset(new OpenCVFivePointTri(&pipe)); else tr.reset(new GpuFivePointTri(&pipe));
    tr->triangulate(src, next, Rm, tm);
    for (int a = 0; a < 3; a++) { for (int b = 0; b < 3; b++) R[3 * a + b] = Rm.at<double>(a, b); t[a] = tm.at<double>(a); }
    *scale = pipe.scale;
    for (int p = 0; p < n; p++) {
        f3d_index[p] = -1;
        auto it = src.map.find(fs[p]);
        if (it == src.map.end() || it->second.expired()) continue;
        std::shared_ptr<Feature3D> f = it->second.lock();
        const int k = (int)(std::find(pipe.feats3d.begin(), pipe.feats3d.end(), f) - pipe.feats3d.begin());
        f3d_index[p] = k;
        cv::Point3f q = f->getPoint();
        points[3 * k] = q.x; points[3 * k + 1] = q.y; points[3 * k + 2] = q.z;
    }
    if (next.map.size() != src.map.size()) throw std::runtime_error("src.map / next.map sizes disagree");
    return (int)pipe.feats3d.size();
    REF_CATCH
}
