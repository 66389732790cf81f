/* oracle/surf_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the reference hot path of Accustomer/CUDA-SURF
 * (surf::Surfor::detectAndCompute / match). Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library; the product
 * (cuda-surf_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED against outputs of the reference itself run on a B200
 * (oracle/_ref/libsurfref.so built from /root/reference by oracle/Makefile; vectors committed
 * under tests/golden/ by tests/golden/make_golden.py). The reference ships no tests or golden
 * vectors of its own.
 */
#ifndef SURF_ORACLE_H
#define SURF_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OR_MAX_SCALE 8  /* /root/reference/surfd.h:9  */
#define OR_MAX_OCTAVE 8 /* /root/reference/surfd.h:10 */

/* Layout of surf::SurfPoint, /root/reference/surf_structures.h:7-31 (48 bytes). */
typedef struct {
    float x, y, scale;
    int o;
    float strength;
    int laplace;
    float ori, score;
    int match;
    float match_x, match_y, ambiguity;
} or_point;

/* surf::SurfParam as derived by Surfor::init, /root/reference/surf.cpp:60-80. */
typedef struct {
    float thresh;
    int init_lobe, doubled, max_scale, noctaves, sampling;
    float divisor;
    int upright, extend, desc_wsz, mag_factor, orient_size, nfeatures;
} or_params;

/* Per-octave parameter schedule, /root/reference/surf.cpp:240-294 + surfd.cu:2829-2865,3058-3073. */
typedef struct {
    int octave;                /* 1,2,4,...                                  */
    int sw, sh;                /* octave dims (swhps[o].x/.y)                */
    int s0, nl;                /* first computed layer, # computed layers    */
    int l[OR_MAX_SCALE];       /* lobe ("mask_sizes") per computed layer     */
    int delta[OR_MAX_SCALE];   /* sampling*octave                            */
    int b1[OR_MAX_SCALE];      /* border actually computed (borders1)        */
    float norm[OR_MAX_SCALE];  /* (9/l^2)^2                                  */
    int borders[OR_MAX_SCALE]; /* lagged borders[] handed to NMS (d_borders) */
    int mb[OR_MAX_SCALE];      /* maximum_borders per cell layer z           */
    int nmb;
} or_octave;

void or_make_params(or_params* p, int noctaves, float thresh, int doubled, int init_mask_size, int sampling_step,
                    int upright, int extend, int desc_wsz);
int or_make_schedule(const or_params* p, int w, int h, or_octave* sched /*[noctaves]*/);
long long or_resp_floats(const or_params* p, const or_octave* sched);

/* integral: img tight pitch `pitch` bytes; out: (h+1) rows x (w+1) cols, tight. */
void or_upsample2x(const uint8_t* img, int w, int h, int pitch, uint8_t* out /* (2h-2) x (2w-2) */);
void or_integral_doubled(const uint8_t* img, int w, int h, int pitch, int32_t* out /* (2h-1) x (2w-1) */);
void or_integral(const uint8_t* img, int w, int h, int pitch, int32_t* out);
/* Hessian: integral tight (w+1)x(h+1); resp: per octave max_scale layers of sw*sh, concatenated. */
void or_hessian(const or_params* p, const or_octave* sched, const int32_t* integral, int w, int h, float* resp);
/* NMS + refine + makePoint. Returns number of keypoints (<= max_pts), scan order o,z,y,x. */
int or_find_keypoints(const or_params* p, const or_octave* sched, const int32_t* integral, int w, int h,
                      const float* resp, or_point* pts, int max_pts);
void or_orientation(const or_params* p, const int32_t* integral, int w, int h, or_point* pts, int n);
/* descriptors [n][nfeatures]; normalised when normalise!=0. */
void or_describe(const or_params* p, const int32_t* integral, int w, int h, const or_point* pts, int n, float* desc,
                 int normalise);
/* brute-force correlation matching with the reference's 8-group top-2 merge. */
void or_match(or_point* pts1, int n1, const float* f1, const or_point* pts2, int n2, const float* f2, int nfeatures);

/* whole frame: integral -> hessian -> keypoints -> (orientation) -> descriptors. Returns n. */
int or_detect_and_compute(const or_params* p, const uint8_t* img, int w, int h, int pitch, or_point* pts, int max_pts,
                          float* desc /*nullable*/);
/* cpu_baseline helper: nframes frames processed with `threads` OpenMP workers; returns seconds. */
double or_time_frames(const or_params* p, const uint8_t* imgs, long long frame_stride, int nframes, int w, int h,
                      int pitch, int max_pts, int threads, long long* total_pts);

#ifdef __cplusplus
}
#endif
#endif
