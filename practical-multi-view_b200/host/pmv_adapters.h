// pmv_adapters.h -- header-only C++11 adapters that put libpmv_cuda.so behind the reference's plugin
// interfaces, so OdometryPipeline keeps working with five edited lines (OdometryPipeline.cpp:68-72):
//
//     extractor = new GpuGoodFeatureExtractor();      // was OpenCVGoodFeatureExtractor
//     matcher   = new GpuLucasKanadeFM();             // was OpenCVLucasKanadeFM
//     ba        = new GpuBundleAdjustment(this);      // was CeresBundleAdjustment
//     pnpsolver = new GpuEPnPSolver(this);            // was OpenCVEPnPSolver
//     triangulator = new GpuFivePointTri(this);      // was OpenCVFivePointTri
//
// Each class mirrors the reference class it replaces -- same members, defaults, Feature fields it
// fills, iteration order over Frame::map -- and only swaps the OpenCV / Ceres call for the C-ABI
// call (include/pmv_cuda.h).  Compile with the reference's include/ directory, OpenCV >= 3.3 headers
// and -lpmv_cuda.  No Ceres needed any more.  Errors surface as std::runtime_error (the reference lets
// cv::Exception propagate the same way).
#ifndef PMV_ADAPTERS_H
#define PMV_ADAPTERS_H

#include <algorithm>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include <opencv2/calib3d.hpp>
#include <opencv2/core.hpp>

#include "BaseFeatureExtractor.h"
#include "BaseFeatureMatcher.h"
#include "BaseOptimizer.h"
#include "BasePnPSolver.h"
#include "BaseTriangulator.h"
#include "OdometryPipeline.h"
#include "pmv_cuda.h"

namespace pmv {

// One context per adapter object: the matcher / extractor run on the producer thread, the optimizer
// on the consumer thread (OdometryPipeline.cpp:210-245), so handles must not be shared.
class Handle {
public:
    explicit Handle(int device = 0) : ctx_(pmv_create(device))
    {
        if (!ctx_) throw std::runtime_error("pmv_create failed: no CUDA device (there is no CPU fallback)");
    }
    ~Handle() { pmv_destroy(ctx_); }
    Handle(const Handle&) = delete;
    Handle& operator=(const Handle&) = delete;
    pmv_ctx* get() const { return ctx_; }
    void check(int rc, const char* what) const
    {
        if (rc != PMV_OK) throw std::runtime_error(std::string(what) + ": " + pmv_last_error(ctx_));
    }
private:
    pmv_ctx* ctx_;
};

// Parent image and ROI of a (possibly non-contiguous) CV_8UC1 view, as Frame::regionOfInterest
// produces them (Frame.cpp:95-117): base pointer of the parent, its size, the view's rectangle.
struct ViewGeometry {
    const unsigned char* base;
    int full_rows, full_cols, step, x, y, w, h;
    explicit ViewGeometry(const cv::Mat& m)
    {
        cv::Size whole; cv::Point ofs;
        m.locateROI(whole, ofs);
        step = (int)m.step;
        base = m.data - (size_t)ofs.y * m.step - ofs.x;
        full_rows = whole.height; full_cols = whole.width;
        x = ofs.x; y = ofs.y; w = m.cols; h = m.rows;
    }
};

}  // namespace pmv

// ---- drop-in for OpenCVLucasKanadeFM (OpenCVLucasKanadeFM.h / .cpp:5-32) -------------------------
class GpuLucasKanadeFM : public BaseFeatureMatcher
{
private:
    int win_size = 32;   // OpenCVLucasKanadeFM.h:9
    int pyr_size = 4;    // OpenCVLucasKanadeFM.h:10
    pmv::Handle gpu;

public:
    virtual fmap matchFeatures(Frame& src, Frame& next)
    {
        fmap correspondences;
        std::vector<float> prev_points;                       // (column, row) pairs, Frame::map iteration order
        for (auto const& p : src.map) {
            prev_points.push_back((float)p.first->column);
            prev_points.push_back((float)p.first->row);
        }
        const int n = (int)src.map.size();
        std::vector<float> next_points(2 * (size_t)n), err(n);
        std::vector<unsigned char> status(n);
        if (n > 0) {
            cv::Mat a = src.bw, b = next.bw;
            if (a.step != b.step) { a = a.clone(); b = b.clone(); }   // the ABI takes one row step for both
            gpu.check(pmv_lk_track(gpu.get(), a.data, b.data, a.rows, a.cols, (int)a.step, prev_points.data(), n,
                                   win_size, win_size, pyr_size, 30, 0.01, 0, 1e-4,   // cv defaults the reference relies on
                                   next_points.data(), status.data(), err.data()),
                      "pmv_lk_track");
        }
        int i = 0;
        for (auto const& p : src.map) {
            if (status[i]) {
                // Feature(int column, int row): float -> int truncation, as at OpenCVLucasKanadeFM.cpp:25
                std::shared_ptr<Feature> f = std::make_shared<Feature>(Feature(next_points[2 * i], next_points[2 * i + 1]));
                next.map[f] = p.second;
                correspondences[p.first] = f;
            }
            i++;
        }
        return correspondences;
    }
};

// ---- drop-in for OpenCVGoodFeatureExtractor (.h / .cpp:4-21): the extractor the pipeline instantiates ----
class GpuGoodFeatureExtractor : public BaseFeatureExtractor
{
public:
    double quality = 0.01;       // OpenCVGoodFeatureExtractor.h:9
    double min_distance = 5;     // OpenCVGoodFeatureExtractor.h:11
    GpuGoodFeatureExtractor() {}
    GpuGoodFeatureExtractor(double quality, double min_distance) : quality(quality), min_distance(min_distance) {}

    std::vector<Feature> extractFeatures(Frame& src, int max)
    {
        pmv::ViewGeometry g(src.bw);
        const size_t cap = max > 0 ? (size_t)max : (size_t)g.w * g.h;
        std::vector<float> xy(2 * cap), score(cap);
        int n = 0;
        gpu.check(pmv_gftt(gpu.get(), g.base, g.full_rows, g.full_cols, g.step, g.x, g.y, g.w, g.h, max, quality,
                           min_distance, 3, 3, xy.data(), score.data(), &n), "pmv_gftt");
        std::vector<Feature> feats;
        for (int i = 0; i < n; i++) {
            Feature f;
            f.row = (int)xy[2 * i + 1];
            f.column = (int)xy[2 * i];
            f.detector = Feature::extractor::cv_good;
            f.tracked = true;                                  // score stays 0, like the reference (.cpp:11-19)
            feats.push_back(f);
        }
        return feats;
    }
private:
    pmv::Handle gpu;
};

// ---- drop-in for ShiTomasiFeatureExtractor (.h / .cpp:5-47) ------------------------------------------
class GpuShiTomasiFeatureExtractor : public BaseFeatureExtractor
{
public:
    double quality = 0.4;        // ShiTomasiFeatureExtractor.h:10
    bool signed_quirk = true;    // Frame.cpp:65-67 reads the u8 image through schar*; false = plain u8
    GpuShiTomasiFeatureExtractor() {}

    virtual std::vector<Feature> extractFeatures(Frame& src, int max)
    {
        const cv::Mat& m = src.bw;
        const size_t cap = max > 0 ? (size_t)max : 1;
        std::vector<int> col(cap), row(cap);
        std::vector<double> score(cap);
        int n = 0;
        gpu.check(pmv_shitomasi(gpu.get(), m.data, m.rows, m.cols, (int)m.step, max > 0 ? max : 0, quality,
                                signed_quirk ? 1 : 0, col.data(), row.data(), score.data(), &n), "pmv_shitomasi");
        std::vector<Feature> best_feats;
        for (int i = 0; i < n; i++) {
            Feature f;                                         // tracked stays false, as in the reference (.cpp:24)
            f.row = row[i];
            f.column = col[i];
            f.detector = Feature::extractor::shi_tomasi;
            f.score = score[i];
            best_feats.push_back(f);
        }
        return best_feats;
    }

    // ShiTomasiFeatureExtractor::computeShiTomasiResponse (.cpp:49-75)
    cv::Mat computeShiTomasiResponse(Frame& src)
    {
        const cv::Mat& m = src.bw;
        cv::Mat dst = cv::Mat::zeros(m.size(), CV_64FC1);
        gpu.check(pmv_shitomasi_response(gpu.get(), m.data, m.rows, m.cols, (int)m.step, signed_quirk ? 1 : 0,
                                         dst.ptr<double>(0)), "pmv_shitomasi_response");
        return dst;
    }
private:
    pmv::Handle gpu;
};

// ---- drop-in for OpenCVFASTFeatureExtractor (.h / .cpp:4-22) ------------------------------------------
class GpuFASTFeatureExtractor : public BaseFeatureExtractor
{
public:
    int threshold = 10;          // OpenCVFASTFeatureExtractor.h:10
    bool nonmax = true;          // OpenCVFASTFeatureExtractor.h:11

    virtual std::vector<Feature> extractFeatures(Frame& src, int max)
    {
        const cv::Mat& m = src.bw;
        const size_t cap = max > 0 ? (size_t)max : 1;
        std::vector<int> col(cap), row(cap);
        std::vector<float> score(cap);
        int n = 0;
        gpu.check(pmv_fast(gpu.get(), m.data, m.rows, m.cols, (int)m.step, threshold, nonmax ? 1 : 0, max > 0 ? max : 0,
                           col.data(), row.data(), score.data(), &n, nullptr), "pmv_fast");
        std::vector<Feature> feats;
        for (int i = 0; i < n; i++) {
            Feature f(cv::Point(col[i], row[i]));              // first `max` keypoints in raster order (.cpp:11-20)
            f.score = score[i];
            f.tracked = true;
            feats.push_back(f);
        }
        return feats;
    }
private:
    pmv::Handle gpu;
};

// ---- drop-in for CeresBundleAdjustment (.h / .cpp:5-89) ----------------------------------------------------
class GpuBundleAdjustment : public BaseOptimizer
{
public:
    OdometryPipeline* tracker;
    pmv_ba_summary last_summary;

    GpuBundleAdjustment(OdometryPipeline* tracker) : tracker(tracker) { std::memset(&last_summary, 0, sizeof last_summary); }

    void apply(Frame& f)
    {
        int fn = (int)f.frame + 1;
        int n = std::min(tracker->bundle_size, fn);
        double K[9];
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) K[3 * r + c] = tracker->camera.at<double>(r, c);

        std::map<int, int> pose_slot;                                       // frame index -> pose block
        std::unordered_map<std::shared_ptr<Feature3D>, int> point_slot;    // Feature3D -> point block
        std::vector<std::shared_ptr<Feature3D>> point_ptr;
        std::vector<double> poses, points, obs;
        std::vector<int32_t> cam_idx, pt_idx;

        for (int i = fn - n; i < fn; i++) {
            if (i == 0) continue;                                           // frame 0 is never a parameter (.cpp:22)
            std::shared_ptr<Frame> frame = tracker->frames[i];
            cv::Mat rod = cv::Mat_<double>(3, 1), R_transpose = cv::Mat_<double>(3, 3);
            cv::transpose(tracker->R[i], R_transpose);
            cv::Rodrigues(R_transpose, rod);
            pose_slot[i] = (int)poses.size() / 6;
            for (int k = 0; k < 3; k++) poses.push_back(rod.at<double>(k));
            for (int k = 0; k < 3; k++) poses.push_back(-tracker->t[i].at<double>(k));
            for (auto& p : frame->map) {
                if (p.second.expired()) continue;
                std::shared_ptr<Feature3D> f3d = p.second.lock();
                if (!point_slot.count(f3d)) {
                    cv::Point3f p3f = f3d->getPoint();
                    point_slot[f3d] = (int)point_ptr.size();
                    point_ptr.push_back(f3d);
                    points.push_back(p3f.x); points.push_back(p3f.y); points.push_back(p3f.z);
                }
                obs.push_back((double)p.first->column); obs.push_back((double)p.first->row);   // .cpp:45
                cam_idx.push_back(pose_slot[i]); pt_idx.push_back(point_slot[f3d]);
            }
        }
        const int Nc = (int)poses.size() / 6, Np = (int)point_ptr.size(), No = (int)cam_idx.size();
        if (Nc > 0 && Np > 0 && No > 0) {
            // HuberLoss(1.0), SPARSE_SCHUR, max_num_iterations = ba_iterations (.cpp:50-61)
            gpu.check(pmv_ba_solve(gpu.get(), poses.data(), points.data(), obs.data(), cam_idx.data(), pt_idx.data(),
                                   Nc, Np, No, K, 1.0, tracker->ba_iterations, &last_summary), "pmv_ba_solve");
            if (tracker->verbose)
                std::cout << "GPU bundle adjustment: cost " << last_summary.initial_cost << " -> " << last_summary.final_cost
                          << " in " << last_summary.iterations << " iterations\n";
        }
        // Updating 3D points and camera poses (.cpp:67-88)
        for (auto& ps : pose_slot) {
            const double* tr = &poses[6 * (size_t)ps.second];
            double __t[] = {tr[3], tr[4], tr[5]};
            double __rod[] = {tr[0], tr[1], tr[2]};
            cv::Mat _R = cv::Mat_<double>(3, 3);
            cv::Mat _t = cv::Mat_<double>(3, 1, __t);
            cv::Mat rod = cv::Mat_<double>(3, 1, __rod);
            cv::Rodrigues(rod, _R);
            cv::transpose(_R, _R);
            tracker->R[ps.first] = _R.clone();
            tracker->t[ps.first] = -_t.clone();
        }
        for (int j = 0; j < Np; j++) point_ptr[j]->update(points[3 * j], points[3 * j + 1], points[3 * j + 2]);
    }
private:
    pmv::Handle gpu;
};

// ---- drop-in for OpenCVEPnPSolver (.h / .cpp:4-50) ---------------------------------------------------------
class GpuEPnPSolver : public BasePnPSolver
{
public:
    OdometryPipeline* tracker;

    GpuEPnPSolver(OdometryPipeline* tracker) : tracker(tracker) {}

    void solvePnP(Frame& src, Frame& next, cv::Mat& R_out, cv::Mat& t_out)
    {
        int j = src.frame;
        std::vector<float> obj, img;
        cv::Mat _R_rod;
        cv::Rodrigues(R_out, _R_rod);
        std::vector<std::weak_ptr<Feature3D>> local_feats3d;

        for (auto& p : src.map) {                                          // same walk as .cpp:13-32
            if (p.second.expired()) continue;
            std::shared_ptr<Feature3D> f3d = p.second.lock();
            if (src.feat_corr[p.first].expired()) continue;
            std::shared_ptr<Feature> f = src.feat_corr[p.first].lock();
            next.map[f] = std::weak_ptr<Feature3D>(f3d);
            f3d->transformInv(tracker->R[j], tracker->t[j]);
            cv::Point3f p3f = f3d->getPoint();
            p3f.z *= -1;
            obj.push_back(p3f.x); obj.push_back(p3f.y); obj.push_back(p3f.z);
            cv::Point2f q = f->getPoint();
            img.push_back(q.x); img.push_back(q.y);
            f3d->transform(tracker->R[j], tracker->t[j]);
            local_feats3d.push_back(f3d);
        }
        const int n = (int)local_feats3d.size();
        double K[9], rv[3], tv[3];
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) K[3 * r + c] = tracker->camera.at<double>(r, c);
        for (int k = 0; k < 3; k++) { rv[k] = _R_rod.at<double>(k); tv[k] = t_out.at<double>(k); }
        std::vector<uint8_t> mask(n > 0 ? n : 1, 0);
        int n_inliers = 0;
        // solvePnPRansac(obj, img, camera, Mat(), rod, t, true, 100, 8, .99, inliers)  (.cpp:34-35)
        gpu.check(pmv_pnp_ransac(gpu.get(), obj.data(), img.data(), n, K, rv, tv, 1, 100, 8.f, .99, mask.data(), &n_inliers), "pmv_pnp_ransac");
        for (int k = 0; k < 3; k++) { _R_rod.at<double>(k) = rv[k]; t_out.at<double>(k) = tv[k]; }
        cv::Rodrigues(_R_rod, R_out);

        // Removing RANSAC outliers (.cpp:38-48)
        for (int i = 0; i < n; i++) {
            if (!mask[i]) {
                if (local_feats3d[i].expired()) continue;
                std::shared_ptr<Feature3D> f3d = local_feats3d[i].lock();
                tracker->feats3d.erase(std::find(tracker->feats3d.begin(), tracker->feats3d.end(), f3d));
            }
        }
    }
private:
    pmv::Handle gpu;
};

// ---- BaseTriangulator ------------------------------------------------------------------------------------------------
// Replaces OpenCVFivePointTri (OpenCVFivePointTri.cpp:5-54): findEssentialMat(RANSAC, 0.99, 1) + recoverPose in one
// launch (pmv_five_point_pose), then the reference's own scale / Feature3D bookkeeping.
class GpuFivePointTri : public BaseTriangulator
{
public:
    OdometryPipeline* tracker;

    GpuFivePointTri(OdometryPipeline* tracker) : tracker(tracker) {}

    void triangulate(Frame& src, Frame& next, cv::Mat& R_out, cv::Mat& t_out)
    {
        int j = src.frame;
        std::vector<double> p1, p2;
        std::vector<std::shared_ptr<Feature>> p1_ptr, p2_ptr;

        for (auto& p : src.feat_corr) {                                    // same walk as .cpp:13-23
            if (p.first.expired() || p.second.expired()) continue;
            std::shared_ptr<Feature> fst = p.first.lock();
            std::shared_ptr<Feature> sec = p.second.lock();
            cv::Point a = fst->getPoint(), b = sec->getPoint();            // integer pixel coordinates
            p1.push_back(a.x); p1.push_back(a.y);
            p2.push_back(b.x); p2.push_back(b.y);
            p1_ptr.push_back(fst);
            p2_ptr.push_back(sec);
        }
        const int n = (int)p1_ptr.size();
        double K[9], E[9], R[9], t[3];
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) K[3 * r + c] = tracker->camera.at<double>(r, c);
        std::vector<uint8_t> mask(n > 0 ? n : 1, 0);
        std::vector<double> tri(4 * (size_t)(n > 0 ? n : 1), 0.0);
        int n_inliers = 0, n_good = 0;
        // findEssentialMat(p1, p2, camera, RANSAC, 0.99, 1, mask); recoverPose(E, p1, p2, camera, R, t, HUGE_VAL, mask, tri)  (.cpp:25-27)
        gpu.check(pmv_five_point_pose(gpu.get(), p1.data(), p2.data(), n, K, 0.99, 1.0, 1000, HUGE_VAL, E, R, t, nullptr, mask.data(),
                                      tri.data(), &n_inliers, &n_good), "pmv_five_point_pose");
        if (n_inliers == 0) throw std::runtime_error("pmv_five_point_pose: no essential matrix (cv::recoverPose would throw)");
        R_out.create(3, 3, CV_64FC1); t_out.create(3, 1, CV_64FC1);
        for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) R_out.at<double>(r, c) = R[3 * r + c]; t_out.at<double>(r) = t[r]; }

        cv::Mat dist = tracker->gt_t[j + tracker->init_offset + 1] - tracker->gt_t[j + tracker->init_offset];   // .cpp:29-35
        tracker->scale = std::sqrt(
            std::pow(dist.at<double>(0), 2) +
            std::pow(dist.at<double>(1), 2) +
            std::pow(dist.at<double>(2), 2));

        t_out = tracker->scale * t_out;

        for (int i = 0; i < n; i++) {                                      // .cpp:37-53
            if (!mask[i]) continue;                                        // removing RANSAC outliers
            const double w = tri[3 * (size_t)n + i];
            std::shared_ptr<Feature3D> f3d_ptr = std::make_shared<Feature3D>(
                tracker->scale * tri[i] / w,
                tracker->scale * tri[(size_t)n + i] / w,
                tracker->scale * tri[2 * (size_t)n + i] / w * -1);
            if (f3d_ptr->getPoint().z < 0) {
                f3d_ptr->transform(tracker->R[j], tracker->t[j]);
                tracker->feats3d.push_back(f3d_ptr);
                next.map[p2_ptr[i]] = std::weak_ptr<Feature3D>(f3d_ptr);
                src.map[p1_ptr[i]] = std::weak_ptr<Feature3D>(f3d_ptr);
            }
        }
    }
private:
    pmv::Handle gpu;
};

#endif  // PMV_ADAPTERS_H
