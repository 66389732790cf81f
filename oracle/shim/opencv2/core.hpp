// Minimal stand-in for <opencv2/core.hpp> so that the reference's surf.cpp compiles in an
// OpenCV-free image. surf.cpp uses only CV_PI from OpenCV (/root/reference/surf.cpp:4,86) and
// relies on the real header pulling in <cmath>/<algorithm> (expf, std::min).
// Test infrastructure only (oracle/_ref build).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#ifndef CV_PI
#define CV_PI 3.1415926535897932384626433832795
#endif
