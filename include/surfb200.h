/* surfb200.h -- C-ABI of libsurfb200.so, the B200-native (sm_100a) SURF hot path.
 *
 * Drop-in boundary for the detect / describe / match path of Accustomer/CUDA-SURF. Each entry
 * point names the reference interface it replaces (file:line under /root/reference). Plain
 * pointers and sizes only; no C++ or torch types. All device pointers are CUDA device memory of
 * the context's device. Functions return SB_OK (0) or a negative sb_status; sb_last_error()
 * gives the text. A context is single-threaded; any number of contexts (one per GPU / host
 * thread) may coexist -- there is no module-level state (the reference keeps its parameters in
 * __constant__/__device__ symbols, surfd.cu:13-24, and is not re-entrant).
 *
 * There is no CPU fallback: every compute entry point fails with SB_ERR_CUDA when no sm_100
 * device is usable.
 */
#ifndef SURFB200_H
#define SURFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_VERSION 200

typedef enum {
    SB_OK = 0,
    SB_ERR_INVALID = -1,     /* bad argument */
    SB_ERR_CUDA = -2,        /* CUDA runtime / launch failure, or no usable device */
    SB_ERR_UNSUPPORTED = -3, /* valid reference configuration this library does not build */
    SB_ERR_NOMEM = -4
} sb_status;

/* Bit-for-bit layout of surf::SurfPoint (surf_structures.h:7-31), 48 bytes. */
typedef struct sb_point {
    float x, y;      /* sub-pixel position in image pixels                      */
    float scale;     /* 1.2 * lobe/3                                            */
    int o;           /* octave index (the reference leaves this field unset)    */
    float strength;  /* interpolated det-of-Hessian                             */
    int laplace;     /* sign of the Laplacian, +1 / -1                          */
    float ori;       /* orientation (radians), 0 when upright                   */
    float score;     /* best correlation (after sb_match)                       */
    int match;       /* index of the best match in set 2, -1 if none            */
    float match_x, match_y;
    float ambiguity; /* second / best score                                     */
} sb_point;

/* Arguments of surf::Surfor::init (surf.h:27-29, surf.cpp:60-91) plus capacities. */
typedef struct sb_params {
    int noctaves;       /* _noctaves                                            */
    float thresh;       /* _thresh                                              */
    int doubled;        /* _doubled: run on the 2x up-sampled frame              */
    int init_mask_size; /* _init_mask_size (9 -> lobe 3, 5 layers per octave)   */
    int sampling_step;  /* _sampling_step                                       */
    int upright;        /* _upright                                             */
    int extend;         /* _extend       (SURF-128)                             */
    int desc_wsz;       /* _desc_wsz                                            */
    int width, height;  /* _width, _height: frame size the context is built for */
    int max_pts;        /* keypoint capacity per frame (SurfData.max_pts)       */
    int batch;          /* frames per batched call (scratch capacity), >= 1     */
    int device;         /* CUDA device ordinal                                  */
    int fresh_desc;     /* 1: sb_detect_and_compute cudaMalloc's a NEW descriptor buffer of num_pts*nfeatures floats on
                           every call and overwrites *d_desc_addr, exactly as cuDescribe does (surfd.cu:3262-3266; the
                           caller frees each). 0 (default): a non-NULL *d_desc_addr is reused.                        */
} sb_params;

/* Derived geometry (surf.cpp:374-390), for callers that size buffers. */
typedef struct sb_info {
    int max_scale, nfeatures;
    int iw, ih, ipitch;        /* integral image (w+1, h+1; 2w-1, 2h-1 if doubled; pitch in ints) */
    int sw[8], sh[8], sp[8];   /* per-octave response dims and pitch             */
    long long resp_floats;     /* tight floats per frame: sum max_scale*sw*sh    */
    int kernels_per_frame;     /* launches per detect+describe pass              */
    int cand_capacity;         /* slots of the NMS candidate queue per frame = number of 2x2x2 cells: a cell yields at most
                                  one candidate, so the queue cannot overflow at any threshold                           */
} sb_info;

typedef struct sb_ctx sb_ctx;

/* Surfor::Surfor + Surfor::init + the lazy allocMemory (surf.cpp:42-91, 374-415). */
int sb_create(sb_ctx** out, const sb_params* params);
/* Surfor::~Surfor (surf.cpp:47-57). */
void sb_destroy(sb_ctx* ctx);
/* Text of the last error on this context (ctx may be NULL: last sb_create error of the thread). */
const char* sb_last_error(const sb_ctx* ctx);
int sb_get_info(const sb_ctx* ctx, sb_info* info);

/* Stream contract. The synchronous entry points (sb_detect_and_compute, sb_match, sb_match_filter, sb_describe, the
 * host-buffer batch calls) run on a context-owned BLOCKING stream: like the reference, whose work all runs on the legacy
 * default stream, they are ordered after everything the caller has enqueued on stream 0 (the default stream of torch, of
 * cudaMemcpy, of kernels launched without a stream) and stream-0 work issued afterwards is ordered after them. A caller
 * that produces inputs on another (non-blocking) stream must synchronise that stream first, or use the *_async entry
 * points, which run on the stream they are given and nowhere else.                                                    */

/* Surfor::detectAndCompute (surf.h:36, surf.cpp:205-355). Synchronous.
 *   d_image   device u8, row pitch `pitch` bytes, w x h must equal the context's size
 *   d_points  device array of max_pts points (SurfData.d_data); all detector fields written
 *   h_points  nullable host array (SurfData.h_data); gets x,y,scale,o,strength,laplace(,ori)
 *             of the first *num_pts points, other fields untouched (surf.cpp:335-342)
 *   d_desc_addr  in/out. NULL: no descriptors. *d_desc_addr == NULL: a device buffer of
 *             max_pts*nfeatures floats is cudaMalloc'ed and stored there (caller cudaFree's it,
 *             as main.cpp:275-282 does). *d_desc_addr != NULL: that buffer (>= max_pts*nfeatures
 *             floats) is reused -- the reference allocates a new one every call and leaks the
 *             old (surfd.cu:3264). With sb_params.fresh_desc = 1 the reference's behaviour is kept
 *             literally: *d_desc_addr is overwritten with a new num_pts*nfeatures allocation per call.
 *   want_desc 0 -> detection only (the `desc` flag).                                          */
int sb_detect_and_compute(sb_ctx* ctx, const uint8_t* d_image, int w, int h, int pitch, sb_point* d_points,
                          sb_point* h_points, int max_pts, int* num_pts, float** d_desc_addr, int want_desc);

/* Surfor::match (surf.h:40, surf.cpp:418-428) -> cuFindMaxCorr (surfd.cu:3550-3566).
 * Writes score, match, match_x, match_y, ambiguity of d_pts1[0..n1) (and h_pts1 if non-NULL).
 * Candidates are the first n2 - n2%32 points of set 2, as in the reference (surfd.cu:2569).    */
int sb_match(sb_ctx* ctx, sb_point* d_pts1, sb_point* h_pts1, int n1, const float* d_feat1, const sb_point* d_pts2,
             int n2, const float* d_feat2);

/* The same matching, enqueued on `stream` (used as given) with no host copy and no synchronisation:
 * for pipelines that keep stereo pairs on the device (BASELINE config 5).
 * All matching calls of a context share ONE scratch (split operands, partial top-2): issue them on a single stream (or
 * order the streams yourself). The scratch only grows: a call that needs more than any call before reallocates it
 * (cudaFree + cudaMalloc: a device-wide synchronisation, not allowed under stream capture) -- run the largest problem once
 * before capturing a graph.                                                                                          */
int sb_match_async(sb_ctx* ctx, sb_point* d_pts1, int n1, const float* d_feat1, const sb_point* d_pts2, int n2,
                   const float* d_feat2, void* stream);

/* Surfor::match for ALL stereo pairs of a detect batch in one launch sequence (BASELINE config 5: what main.cpp:246-251
 * does per pair), with the keypoint counts read on the device: nothing travels to the host between
 * sb_detect_batch_async and the matching.
 *   d_points [nframes][pts_stride] sb_point, d_counts [nframes], d_desc [nframes][desc_stride floats]: the outputs of
 *            sb_detect_batch_async (pts_stride = max_pts, desc_stride = max_pts * nfeatures)
 *   d_pairs  device array of npairs (frame1, frame2) index pairs; NULL: pair z is (2z, 2z + 1)
 *   bound    upper bound of the keypoint counts that take part (a frame with more is matched on its first `bound`
 *            points); sizes the scratch: npairs * bound * (6 * nfeatures + 128 * splits) bytes, grown on the first call
 * Per pair the results are those of sb_match (score, match, match_x, match_y, ambiguity of frame1's points).
 * 64- and 128-d descriptors only. Enqueued on `stream`, no synchronisation.                                        */
int sb_match_pairs_async(sb_ctx* ctx, sb_point* d_points, long long pts_stride, const int* d_counts, const float* d_desc,
                         long long desc_stride, int npairs, const int* d_pairs, int bound, void* stream);

/* Consumer-side acceptance of match results (the reference draws every row's best candidate,
 * main.cpp:59-70, and leaves the ratio test to the consumer of SurfPoint::ambiguity). Rows of set 1 with
 * match >= 0 and ambiguity < max_ambiguity are compacted, in row order, into (idx1, idx2, score, ambiguity).
 *   flags  SB_FILTER_LAPLACE: also require equal Laplacian signs (surf_structures.h:13)
 *          SB_FILTER_CROSS:   also require d_pts2[idx2].match == idx1 (set 2 matched back with a second sb_match)
 *   d_pairs  device array of `cap` pairs; h_pairs nullable host copy; *num_pairs the number kept (<= cap).       */
typedef struct sb_pair { int idx1, idx2; float score, ambiguity; } sb_pair;
#define SB_FILTER_LAPLACE 1
#define SB_FILTER_CROSS 2
int sb_match_filter(sb_ctx* ctx, const sb_point* d_pts1, int n1, const sb_point* d_pts2, int n2, float max_ambiguity,
                    int flags, sb_pair* d_pairs, sb_pair* h_pairs, int cap, int* num_pairs);

/* Batched, asynchronous form of detectAndCompute for independent frames (the frame loop of
 * main.cpp:239-245 without a host round trip per frame). nframes <= params.batch.
 *   d_images  frame f at d_images + f*image_stride (bytes), row pitch `pitch`
 *   d_points  [nframes][max_pts]; d_counts [nframes] (clamped to max_pts); d_desc nullable
 *             [nframes][max_pts][nfeatures]
 *   stream    cudaStream_t, used as given (NULL is the CUDA default stream). Returns after
 *             enqueueing; results are ordered on that stream.                                 */
int sb_detect_batch_async(sb_ctx* ctx, const uint8_t* d_images, size_t image_stride, int pitch, int nframes,
                          sb_point* d_points, int* d_counts, float* d_desc, void* stream);
/* End-to-end form with HOST buffers: H2D of the frames, the batch above, D2H of counts, points
 * and descriptors, synchronised on return. h_images tight (pitch == width).                   */
int sb_detect_batch_host(sb_ctx* ctx, const uint8_t* h_images, int nframes, sb_point* h_points, int* h_counts,
                         float* h_desc);
/* The same in two halves, for a caller that streams batches: sb_submit_batch_host enqueues the uploads and kernels of a
 * batch and returns a ticket (at most THREE outstanding); sb_wait_batch_host downloads that batch's counts, points and
 * descriptors and returns when they are in the caller's buffers. With two batches submitted ahead, batch k downloads
 * while batch k+1 computes and batch k+2 uploads. h_images must stay valid until the ticket has been waited for.
 * h_points [nframes][max_pts], h_desc [nframes][max_pts][nfeatures]: entries [0, h_counts[f]) of frame f are results;
 * entries past a frame's own count (up to the largest count of its chunk) are overwritten with unspecified values.    */
int sb_submit_batch_host(sb_ctx* ctx, const uint8_t* h_images, int nframes, int want_desc, int* ticket);
int sb_wait_batch_host(sb_ctx* ctx, int ticket, sb_point* h_points, int* h_counts, float* h_desc);
int sb_sync(sb_ctx* ctx);
/* Same work as sb_detect_batch_async, synchronous, with CUDA events recorded on the launching stream
 * at the stage boundaries: stage_ms[0..3] = integral, Hessian, NMS+refine(+clamp), orientation+describe.
 * Measurement aid for bench.py (per-kernel roofline); not a reference interface.                */
int sb_detect_batch_profile(sb_ctx* ctx, const uint8_t* d_images, size_t image_stride, int pitch, int nframes,
                            sb_point* d_points, int* d_counts, float* d_desc, void* stream, float* stage_ms);

/* Stage access for parity tests (the reference exposes these stages as cuIntegral,
 * cuCalcHessianMulti, cuDescribe in surfd.h:63,104,132). `slot` is the frame slot of the last
 * detect call (0 for sb_detect_and_compute). Outputs are tight host arrays.                   */
int sb_get_integral(sb_ctx* ctx, int slot, int32_t* h_out /* (h+1)*(w+1) */);
int sb_get_response(sb_ctx* ctx, int slot, float* h_out /* sb_info.resp_floats */);
/* (orientation if !upright) + descriptors + normalisation for caller-supplied points on the
 * integral image held in `slot` (cuDescribe, surfd.cu:3251-3325). d_points' ori is updated.   */
int sb_describe(sb_ctx* ctx, int slot, sb_point* d_points, int n, float* d_desc);

/* Workload generator `synth_v1`: deterministic textured frame, host memory. */
int sb_synth_frame(uint8_t* out, int w, int h, int pitch, uint64_t seed, int shift_x, int noise_amp,
                   uint64_t noise_seed);

#ifdef __cplusplus
}
#endif
#endif /* SURFB200_H */
