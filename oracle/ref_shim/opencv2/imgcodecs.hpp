// TEST INFRASTRUCTURE ONLY -- no codecs in the shim: imread returns an empty Mat (frames are built from arrays).
#pragma once
#include "core.hpp"
namespace cv { enum { IMREAD_COLOR = 1 }; inline Mat imread(const String&, int = IMREAD_COLOR) { return Mat(); } }
