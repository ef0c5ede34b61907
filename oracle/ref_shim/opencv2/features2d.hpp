// TEST INFRASTRUCTURE ONLY -- cv::KeyPoint + cv::FAST for OpenCVFASTFeatureExtractor.cpp.
#pragma once
#include "core.hpp"
namespace cv {
struct KeyPoint { Point2f pt; float size, angle, response; int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {} };
void FAST(const Mat& image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
}
