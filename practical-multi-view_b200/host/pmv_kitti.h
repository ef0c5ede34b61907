// pmv_kitti.h -- the KITTI odometry wire formats the pipeline reads and the error report it writes,
// host side, no OpenCV needed.  Mirrors OdometryPipeline::parsePoses (OdometryPipeline.cpp:525-593), ::parseCalibration
// (:595-653), ::split (:497-520), ::standardDeviation (:657-669) and the error report of ::run (:272-296) -- including
// what the reference's parsing does with odd input (tokens are split on single spaces and empty ones dropped; a token
// that is not a number parses as 0; the calibration line loses its last token, which the 3x3 camera never needs).
// The same functions are exported through the C ABI (pmv_kitti_* in include/pmv_cuda.h).
#ifndef PMV_KITTI_H
#define PMV_KITTI_H

#include <algorithm>
#include <array>
#include <cmath>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

namespace pmv {
namespace kitti {

// OdometryPipeline::split with the default delimiter " "
inline std::vector<std::string> split(const std::string& str, const std::string& delim = " ")
{
    std::vector<std::string> tokens;
    std::size_t prev = 0, pos = 0;
    do {
        pos = str.find(delim, prev);
        if (pos == std::string::npos) pos = str.length();
        std::string token = str.substr(prev, pos - prev);
        if (!token.empty()) tokens.push_back(token);
        prev = pos + delim.length();
    } while (pos < str.length() && prev < str.length());
    return tokens;
}

inline double to_double(const std::string& s)
{
    std::stringstream ss(s);
    double j = 0.0;
    ss >> j;
    return j;
}

// poses file: one "r00 r01 r02 tx r10 r11 r12 ty r20 r21 r22 tz" line per frame, at most `stop` lines.
// R: row-major 3x3 per frame, t: 3 per frame.  Returns false when the file cannot be opened (the reference throws).
inline bool parse_poses(const std::string& filename, int stop, std::vector<std::array<double, 9>>& R, std::vector<std::array<double, 3>>& t)
{
    std::ifstream f(filename);
    if (!f) return false;
    std::string line;
    int k = 0;
    while (std::getline(f, line) && k < stop) {
        k++;
        std::array<double, 9> Rk{};
        std::array<double, 3> tk{};
        int i = 0;
        for (const std::string& tok : split(line)) {
            const double j = to_double(tok);
            if (i < 12) {
                if (i % 4 == 3) tk[i / 4] = j; else Rk[3 * (i / 4) + i % 4] = j;
            }
            i++;
        }
        R.push_back(Rk);
        t.push_back(tk);
    }
    return true;
}

// calib.txt: line `num_calib` is "Pn: p00 p01 p02 p03 p10 ... p23"; the camera matrix is its left 3x3.
// Only tokens FOLLOWED by a space are consumed (the reference erases up to each space), K keeps its previous
// content where the line is short.  Returns false when the file cannot be opened.
inline bool parse_calibration(const std::string& filename, int num_calib, double K[9])
{
    std::ifstream f(filename);
    if (!f) return false;
    std::string calib;
    int i = 0;
    while (std::getline(f, calib)) {
        if (i == num_calib) {
            int k = 0;
            std::size_t pos;
            while ((pos = calib.find(" ")) != std::string::npos) {
                const double j = to_double(calib.substr(0, pos));
                calib.erase(0, pos + 1);
                if (k >= 1 && k <= 11 && k % 4 != 0) K[3 * ((k - 1) / 4) + (k - 1) % 4] = j;
                k++;
            }
        }
        i++;
    }
    return true;
}

inline double standard_deviation(const std::vector<double>& val)
{
    double avg = 0, sd = 0;
    for (double v : val) avg += v;
    avg /= val.size();
    for (double v : val) sd += std::pow(v - avg, 2);
    return std::sqrt(sd / (val.size() - 1));
}

struct ErrorReport {
    double R_total = 0, R_min = 0, R_max = 0, R_std = 0, t_total = 0, t_min = 0, t_max = 0, t_std = 0;
    std::vector<double> errors_R, errors_t;
};

// The error loop of OdometryPipeline::run (:272-288) on n estimated poses (index 0 = the initial identity).  The
// reference flips gt_t[i + off].z and gt_R[i + off](2,0), (0,2) IN PLACE (cv::Mat copies share storage) and then
// compares R[i] with gt_R[i] -- without the offset; both are reproduced on private copies of the ground truth.
inline ErrorReport error_report(const std::vector<std::array<double, 9>>& R, const std::vector<std::array<double, 3>>& t,
                                std::vector<std::array<double, 9>> gt_R, std::vector<std::array<double, 3>> gt_t, int init_offset)
{
    ErrorReport rep;
    for (std::size_t i = 1; i < t.size(); i++) {
        std::array<double, 3>& g = gt_t[i + init_offset];
        g[2] *= -1;
        std::array<double, 9>& G = gt_R[i + init_offset];
        G[6] *= -1;
        G[2] *= -1;
        double st = 0, sr = 0;
        for (int k = 0; k < 3; k++) st += (t[i][k] - g[k]) * (t[i][k] - g[k]);
        for (int k = 0; k < 9; k++) sr += (R[i][k] - gt_R[i][k]) * (R[i][k] - gt_R[i][k]);
        rep.errors_t.push_back(std::sqrt(st));
        rep.errors_R.push_back(std::sqrt(sr));
        rep.t_total += std::sqrt(st);
        rep.R_total += std::sqrt(sr);
    }
    if (!rep.errors_R.empty()) {
        rep.R_min = *std::min_element(rep.errors_R.begin(), rep.errors_R.end());
        rep.R_max = *std::max_element(rep.errors_R.begin(), rep.errors_R.end());
        rep.R_std = standard_deviation(rep.errors_R);
        rep.t_min = *std::min_element(rep.errors_t.begin(), rep.errors_t.end());
        rep.t_max = *std::max_element(rep.errors_t.begin(), rep.errors_t.end());
        rep.t_std = standard_deviation(rep.errors_t);
    }
    return rep;
}

// the text of error_path (:289-300), default ostream formatting like the reference
inline std::string format_report(double runtime, const ErrorReport& r)
{
    std::ostringstream f;
    f << "Runtime: " << runtime << std::endl;
    f << "R total: " << r.R_total << std::endl;
    f << "R min: " << r.R_min << std::endl;
    f << "R max: " << r.R_max << std::endl;
    f << "R std: " << r.R_std << std::endl;
    f << "t total: " << r.t_total << std::endl;
    f << "t min: " << r.t_min << std::endl;
    f << "t max: " << r.t_max << std::endl;
    f << "t std: " << r.t_std << std::endl;
    return f.str();
}

}  // namespace kitti
}  // namespace pmv
#endif  // PMV_KITTI_H
