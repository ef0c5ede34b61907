// TEST INFRASTRUCTURE ONLY -- functional stand-in for the slice of <opencv2/calib3d.hpp> the reference uses.
#pragma once
#include "core.hpp"
namespace cv {
void Rodrigues(const Mat& src, OutputArray dst);   // 3x3 <-> 3x1 / 1x3, CV_64F
bool solvePnPRansac(const std::vector<Point3f>& obj, const std::vector<Point2f>& img, const Mat& K, const Mat& dist,
                    OutputArray rvec, OutputArray tvec, bool useExtrinsicGuess, int iterationsCount, float reprojectionError,
                    double confidence, std::vector<int>& inliers);
enum { LMEDS = 4, RANSAC = 8 };
Mat findEssentialMat(const std::vector<Point>& p1, const std::vector<Point>& p2, const Mat& K, int method, double prob, double threshold,
                     OutputArray mask);
int recoverPose(const Mat& E, const std::vector<Point>& p1, const std::vector<Point>& p2, const Mat& K, OutputArray R, OutputArray t,
                double distanceThresh, OutputArray mask, OutputArray triangulatedPoints);
}
