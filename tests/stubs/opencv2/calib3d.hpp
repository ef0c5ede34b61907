#pragma once
#include "core.hpp"
namespace cv { void Rodrigues(const Mat&, Mat&); }
