// TEST INFRASTRUCTURE ONLY -- cv::calcOpticalFlowPyrLK for OpenCVLucasKanadeFM.cpp.
#pragma once
#include "../core.hpp"
namespace cv {
void calcOpticalFlowPyrLK(const Mat& prevImg, const Mat& nextImg, const std::vector<Point2f>& prevPts,
                          std::vector<Point2f>& nextPts, std::vector<uchar>& status, std::vector<float>& err,
                          Size winSize = Size(21, 21), int maxLevel = 3,
                          TermCriteria criteria = TermCriteria(TermCriteria::COUNT + TermCriteria::EPS, 30, 0.01),
                          int flags = 0, double minEigThreshold = 1e-4);
}
