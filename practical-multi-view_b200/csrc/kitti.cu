// kitti.cu -- C-ABI face of host/pmv_kitti.h (KITTI wire formats + error report).  Host code only:
// the formats are text files read once per run; they are part of the boundary so that a pipeline built on the C ABI
// reads the same calib.txt / poses and writes the same report as OdometryPipeline.cpp:525-653, 272-300.
#include <cstring>

#include "../host/pmv_kitti.h"
#include "common.cuh"

extern "C" {

PMV_API int pmv_kitti_parse_poses(const char *path, int stop, double *R, double *t, int capacity, int *n)
{
    if (!path || !n || capacity < 0) return PMV_ERR_INVALID;
    std::vector<std::array<double, 9>> Rv;
    std::vector<std::array<double, 3>> tv;
    if (!pmv::kitti::parse_poses(path, stop, Rv, tv)) return PMV_ERR_INVALID;   // "Unable to open pose file"
    *n = (int)Rv.size();
    const int m = *n < capacity ? *n : capacity;
    for (int i = 0; i < m; i++) {
        if (R) std::memcpy(R + 9 * (size_t)i, Rv[i].data(), 9 * sizeof(double));
        if (t) std::memcpy(t + 3 * (size_t)i, tv[i].data(), 3 * sizeof(double));
    }
    return PMV_OK;
}

PMV_API int pmv_kitti_parse_calibration(const char *path, int num_calib, double K[9])
{
    if (!path || !K) return PMV_ERR_INVALID;
    return pmv::kitti::parse_calibration(path, num_calib, K) ? PMV_OK : PMV_ERR_INVALID;   // "Unable to open calibration file"
}

PMV_API int pmv_kitti_error_report(const double *R, const double *t, int n, const double *gt_R, const double *gt_t, int n_gt,
                                   int init_offset, double runtime, double stats[8], char *text, int text_capacity)
{
    if (!R || !t || !gt_R || !gt_t || n < 1 || init_offset < 0 || n - 1 + init_offset >= n_gt) return PMV_ERR_INVALID;
    std::vector<std::array<double, 9>> Rv(n), Gv(n_gt);
    std::vector<std::array<double, 3>> tv(n), gv(n_gt);
    for (int i = 0; i < n; i++) { std::memcpy(Rv[i].data(), R + 9 * (size_t)i, 72); std::memcpy(tv[i].data(), t + 3 * (size_t)i, 24); }
    for (int i = 0; i < n_gt; i++) { std::memcpy(Gv[i].data(), gt_R + 9 * (size_t)i, 72); std::memcpy(gv[i].data(), gt_t + 3 * (size_t)i, 24); }
    const pmv::kitti::ErrorReport r = pmv::kitti::error_report(Rv, tv, Gv, gv, init_offset);
    if (stats) { const double s[8] = {r.R_total, r.R_min, r.R_max, r.R_std, r.t_total, r.t_min, r.t_max, r.t_std}; std::memcpy(stats, s, sizeof s); }
    if (text && text_capacity > 0) {
        const std::string s = pmv::kitti::format_report(runtime, r);
        std::strncpy(text, s.c_str(), text_capacity - 1);
        text[text_capacity - 1] = 0;
    }
    return PMV_OK;
}

}  // extern "C"
