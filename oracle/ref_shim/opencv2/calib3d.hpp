// TEST INFRASTRUCTURE ONLY -- functional stand-in for the slice of <opencv2/calib3d.hpp> the reference uses.
#pragma once
#include "core.hpp"
namespace cv {
void Rodrigues(const Mat& src, OutputArray dst);   // 3x3 <-> 3x1 / 1x3, CV_64F
bool solvePnPRansac(const std::vector<Point3f>& obj, const std::vector<Point2f>& img, const Mat& K, const Mat& dist,
                    OutputArray rvec, OutputArray tvec, bool useExtrinsicGuess, int iterationsCount, float reprojectionError,
                    double confidence, std::vector<int>& inliers);
}
