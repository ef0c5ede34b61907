// TEST INFRASTRUCTURE ONLY -- functional stand-in for the slice of <opencv2/imgproc.hpp> the reference uses.
#pragma once
#include "core.hpp"
#define CV_THRESH_BINARY 0
namespace cv {
enum { THRESH_BINARY = 0 };
enum { COLOR_BGR2GRAY = 6 };
void blur(const Mat& src, OutputArray dst, Size ksize);                                     // hook or 9-term sums
void threshold(const Mat& src, OutputArray dst, double thresh, double maxval, int type);   // CV_64FC1, THRESH_BINARY
void cvtColor(const Mat& src, OutputArray dst, int code);                                   // BGR2GRAY, 8-bit fixed point
void goodFeaturesToTrack(const Mat& image, std::vector<Point2f>& corners, int maxCorners, double qualityLevel,
                         double minDistance, const Mat& mask = Mat(), int blockSize = 3, int gradientSize = 3,
                         bool useHarrisDetector = false, double k = 0.04);
}
