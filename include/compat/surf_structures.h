// surf_structures.h -- drop-in for /root/reference/surf_structures.h (same names, same layouts).
// SurfPoint is bit-for-bit sb_point (48 bytes, surf_structures.h:7-31); SurfData is the same four
// fields (surf_structures.h:35-41). SurfParam (surf_structures.h:44-72) is kept for source
// compatibility of callers that name it; the library derives its own copy in sb_create.
#pragma once
#include "surfb200.h"

namespace surf {

struct SurfPoint {
    float x = -1, y = -1;   // position
    float scale = 1;        // detected scale
    int o = 0;              // octave (written by this implementation; garbage in the reference)
    float strength = 0;     // interpolated det(Hessian)
    int laplace = 1;        // sign of the Laplacian
    float ori = 0;          // orientation
    float score = 0;        // match score
    int match = -1;         // matched index
    float match_x = 0, match_y = 0;
    float ambiguity = 0;    // second / best
};
static_assert(sizeof(SurfPoint) == sizeof(sb_point) && sizeof(SurfPoint) == 48, "SurfPoint layout must match the reference");

struct SurfData {
    int num_pts;        // number of available points
    int max_pts;        // number of allocated points
    SurfPoint* h_data;  // host
    SurfPoint* d_data;  // device
};

struct SurfParam {
    float thresh;
    int init_lobe;
    bool doubled;
    int max_scale, noctaves, sampling;
    float divisor;
    bool upright, extend;
    int desc_wsz, mag_factor, orient_size, nfeatures;
};

}  // namespace surf
