// common.cuh -- parameter blocks and device helpers shared by the sm_100a SURF kernels.
//
// Every kernel gets the whole per-context parameter block by value as a __grid_constant__
// argument (constant bank, dynamically indexable), so there is no module-level device state:
// contexts are independent (the reference keeps this in __constant__ symbols re-uploaded with
// ~22 cudaMemcpyToSymbol per frame, surfd.cu:13-24, 2868-2871, 3072-3073).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "surfb200.h"

namespace sb {

constexpr int kMaxScale = 8;   // surfd.h:9
constexpr int kMaxOctave = 8;  // surfd.h:10
constexpr int kNBin = 72;      // surfd.h:11
constexpr int kHwn = 6;        // surfd.h:14
constexpr int kHessRows = 32;  // output rows per tile of the gather Hessian kernel (hessian.cu); 32 columns

// One octave of the scale space (surf.cpp:240-294, surfd.cu:2844-2865, 3062-3073).
struct OctaveP {
    int sw, sh, sp;          // response dims and row pitch (floats)
    int osz;                 // sh * sp : floats per layer
    int s0, nl;              // first computed layer, number of computed layers
    int octave;              // 1, 2, 4, ...
    int delta;               // sampling * octave
    long long resp_off;      // float offset of layer 0 inside one frame's response buffer
    int l[kMaxScale];        // lobe per computed layer i (layer index s0+i)
    int b1[kMaxScale];       // border actually computed per computed layer
    float norm[kMaxScale];   // (9/l^2)^2
    int borders[kMaxScale];  // lagged borders per layer index (d_borders of the reference)
    int mb[4];               // NMS border per cell layer z (maximum_borders)
    int nmb;
    int hess_tile0, hess_tx, hess_ty;  // first linear tile id / tile grid of this octave (Hessian)
    int nms_tile0, nms_tx, nms_ty;     // same for the NMS cell grid (per z)
    float inv_hess_tx, inv_nms_tx;     // 1 / hess_tx, 1 / nms_tx: tile decode without integer division (tile ids < 2^23)
};

struct PipeP {
    // frame and buffers
    int w, h;                 // image size
    int iw, ih, ip;           // integral dims, pitch in ints
    long long istride;        // ints per frame slot of the integral buffer (incl. 2 guard rows)
    long long rstride;        // floats per frame slot of the response buffer
    int nbands, nchunks;      // integral tiling: bands of kBandRows rows, chunks of 256 output columns
    int band_rows;
    // SurfParam (surf_structures.h:44-72)
    float thresh, divisor;
    int init_lobe, max_scale, noctaves, sampling;
    int upright, extend, desc_wsz, mag_factor, orient_size, nfeatures;
    int doubled;              // 1: w, h, iw, ih describe the 2x up-sampled frame (surf.cpp:69-72, 234-235)
    int max_pts;
    int hess_tiles, nms_tiles;  // total linear tiles per frame
    int nms_cells;              // 2x2x2 NMS cells per frame = capacity of the candidate queue (it cannot overflow)
    OctaveP oct[kMaxOctave];
    // exp tables of Surfor::initLut (surf.cpp:358-371) and the angle bins of surf.cpp:83-90
    float lut1[83];
    float lut2[40];
    float bins[kNBin];
};
static_assert(sizeof(PipeP) <= 4000, "PipeP must fit the 4 KB kernel parameter space");

// Inclusive pixel-box sum x in [xlo,xhi], y in [ylo,yhi] from the padded integral image
// (I[y+1][x+1] = sum over pixels <= (x,y)); same four corners as getSum, surfd.cu:334-343.
__device__ __forceinline__ int box_sum(const int* __restrict__ I, int ip, int xlo, int xhi, int ylo, int yhi) {
    const int r1 = (yhi + 1) * ip, r0 = ylo * ip;
    return __ldg(I + r1 + xhi + 1) + __ldg(I + r0 + xlo) - __ldg(I + r0 + xhi + 1) - __ldg(I + r1 + xlo);
}

// Haar wavelets of surfd.cu:1171-1182: upper-minus-lower and right-minus-left halves of the
// (2s+1)^2 window centred at (x,y).
__device__ __forceinline__ int haar_y(const int* __restrict__ I, int ip, int x, int y, int s) {
    return box_sum(I, ip, x - s, x + s, y - s, y) - box_sum(I, ip, x - s, x + s, y, y + s);
}
__device__ __forceinline__ int haar_x(const int* __restrict__ I, int ip, int x, int y, int s) {
    return box_sum(I, ip, x, x + s, y - s, y + s) - box_sum(I, ip, x - s, x, y - s, y + s);
}

// Both Haar responses of the (2s+1)^2 window at (x, y) from the 12 distinct corners the two box pairs share (haar_x and
// haar_y above read 16), with ONE 32-bit element index per row and 64-bit row pointers made opaque: written through
// box_sum the compiler spent ~70 instructions of 64-bit carry chains on the 16 addresses of a sample (ncu source view of the
// rotated descriptor kernel, round 2). Integer arithmetic, so the values are those of haar_x / haar_y exactly.
__device__ __forceinline__ void haar_xy(const int* __restrict__ I, int ip, int x, int y, int s, int& hx, int& hy) {
    const int base = y * ip + x, sip = s * ip;
    const int* pm = I + (base - sip);       // row y - s
    const int* pz = I + base;               // row y (row y + 1 is ip further)
    const int* pq = I + (base + sip + ip);  // row y + s + 1
    asm volatile("" : "+l"(pm), "+l"(pz), "+l"(pq));
    // columns A = x - s, B = x, C = x + 1, D = x + s + 1
    const int mA = __ldg(pm - s), mB = __ldg(pm), mC = __ldg(pm + 1), mD = __ldg(pm + s + 1);
    const int qA = __ldg(pq - s), qB = __ldg(pq), qC = __ldg(pq + 1), qD = __ldg(pq + s + 1);
    const int zA = __ldg(pz - s), zD = __ldg(pz + s + 1), uA = __ldg(pz + (ip - s)), uD = __ldg(pz + (ip + s + 1));
    hx = (qD + mB - mD - qB) - (qC + mA - mC - qA);
    hy = (zD - zA) + (uD - uA) - (mD - mA) - (qD - qA);
}

// The geometry of one keypoint's sampling lattice (surfd.cu:1578-1590), shared by both descriptor kernels.
struct KpGeom {
    float fx, fy, spacing;
    int ixc, iyc, step, S, R, side, e;
};
// doubled: the descriptor is sampled on the 2x image at (2x, 2y) with 3.3*scale (surfd.cu:1581-1592)
__device__ __forceinline__ KpGeom kp_geom(float x, float y, float scale, int W, int mag_factor, int doubled) {
    KpGeom g;
    if (doubled) { x = __fadd_rn(x, x); y = __fadd_rn(y, y); }
    const float sc = __fmul_rn(doubled ? 3.3f : 1.65f, scale);
    g.step = max(__float2int_rn(__fmul_rn(sc, 0.5f)), 1);
    g.ixc = __float2int_rn(x);
    g.iyc = __float2int_rn(y);
    g.fx = __fsub_rn(x, __int2float_rn(g.ixc));
    g.fy = __fsub_rn(y, __int2float_rn(g.iyc));
    g.spacing = __fmul_rn(sc, __int2float_rn(mag_factor));
    g.S = __float2int_rz(sc);
    g.R = __float2int_rn(__fdiv_rn(__fmul_rn(__fmul_rn(g.spacing, __int2float_rn(W + 1)), 0.5f), __int2float_rn(g.step)));
    g.side = 2 * g.R + 1;
    g.e = 2 * g.step - g.S;
    return g;
}
// q = a / b for 0 <= a < 2^23 with inv = 1.f / b: exact (the +0.5 keeps the product away from the integer boundary)
__device__ __forceinline__ int div_small(int a, float inv) { return __float2int_rz(__fmul_rn(__int2float_rn(a) + 0.5f, inv)); }

// Column-phase layout of the second integral copy that the gather Hessian reads: within a row, column X sits at
// (X % 8) * (pitch / 8) + X / 8. Samples of octave o are 2^(o+1) pixels apart, so the 32 lanes of a gather read 4-byte
// words 8..128 bytes apart in the row-major image (16-32 sectors per request); in this layout the same lanes read
// consecutive words of one or two planes (4-8 sectors up to octave 2).
__device__ __forceinline__ int phase_col(int X, int pitch) { return (X & 7) * (pitch >> 3) + (X >> 3); }

// Programmatic dependent launch between the kernels of a frame: a kernel launched with launch_dep() may be placed on the SMs
// while its predecessor in the stream is still draining; it executes pdl_wait() before its first memory access, which blocks
// until the predecessor has completed and its writes are visible, so the data flow is that of plain stream order -- only the
// launch latency and the ramp-up of every kernel overlap the tail of the one before (8 launches per frame: it matters for
// the single-frame latency, not for batches). pdl_wait() is a no-op in a kernel launched without the attribute.
// Opt-in (SURFB200_PDL=1, read once): on the B200 it made the single frame SLOWER (ctx.cpp: pdl_enabled), so by default every
// kernel is launched plainly and pdl_wait() does nothing.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_dep(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// launchers (one translation unit per stage)
cudaError_t launch_upsample2x(const uint8_t* d_src, size_t src_stride, int src_pitch, int w, int h, uint8_t* d_dst,
                              size_t dst_stride, int dst_pitch, int nframes, cudaStream_t st);
cudaError_t launch_integral(const PipeP& P, const uint8_t* d_images, size_t image_stride, int pitch, int nframes,
                            int* d_integral, int* d_integral_ph, int* d_colsum, int* d_rowsum, int* d_tilesum,
                            int* d_counts /* one per frame, zeroed by the first kernel; may be null */, cudaStream_t st);
cudaError_t launch_hessian(const PipeP& P, int nframes, const int* d_integral, const int* d_integral_ph, float* d_resp,
                           cudaStream_t st);
// d_done: one zeroed unsigned per frame (self re-arming); d_work, d_work_orient, d_cls_cnt: re-armed by the last refine block
cudaError_t launch_nms(const PipeP& P, int nframes, const int* d_integral, const float* d_resp, sb_point* d_points,
                       int* d_counts, unsigned* d_cand, int* d_cand_count, int cand_cap, unsigned* d_done, int* d_work,
                       int* d_work_orient, int* d_cls_cnt /* 4 ints per frame, may be null */, cudaStream_t st);
// What the TMA descriptor path (describe_tma.cu) needs besides the integral image: the tensor maps over the context's
// integral buffer, the per-slot keypoint class lists [batch][2][max_pts] and their counters [batch][4] = {keypoints on the
// TMA path, on the gather path, work counter of either kernel}, zero on entry (the last nms_refine block / sb_describe re-arm them).
struct DescAux {
    const void* maps = nullptr;   // device array of CUtensorMap (128 B each); nullptr: gather kernel only
    int* cls_idx = nullptr;
    int* cls_cnt = nullptr;
    int slot0 = 0;                // first frame slot of this launch inside the context's buffers
    // called (if set) right before the last descriptor kernel is launched -- after the orientation pass of the rotated path --
    // when the keypoints are final: the single-frame call starts their copy to the host there
    cudaError_t (*before_last)(void*) = nullptr;
    void* before_last_arg = nullptr;
};
cudaError_t launch_describe(const PipeP& P, int nframes, const int* d_integral, sb_point* d_points, long long pts_stride,
                            const int* d_counts, int fixed_count, float* d_desc, long long desc_stride, int sm_count,
                            int* d_work, int* d_work_orient /* one zeroed int per frame each */, const DescAux& aux, cudaStream_t st);
cudaError_t build_describe_maps(const PipeP& P, const int* d_integral, int batch, void** d_maps);
bool describe_tma_applies(const PipeP& P);
cudaError_t launch_describe_tma(const PipeP& P, int nframes, const DescAux& aux, sb_point* d_points, long long pts_stride,
                                const int* d_counts, int fixed_count, float* d_desc, long long desc_stride, int sm_count,
                                cudaStream_t st);
// grow-only device scratch of the matcher (split-bf16 operands, per-split group top-2), owned by the context
struct MatchScratch {
    void* a = nullptr; void* b = nullptr; void* part = nullptr;
    size_t cap_a = 0, cap_b = 0, cap_part = 0;
    // the two tensor maps of the last call (CUtensorMap is 128 bytes) and what they were encoded for: a caller that matches
    // sets of the same padded sizes again -- a video stream -- does not pay cuTensorMapEncodeTiled twice per call
    alignas(64) unsigned char map_a[128], map_b[128];
    const void* map_a_base = nullptr; const void* map_b_base = nullptr;
    int map_a_rows = 0, map_b_rows = 0, map_nf = 0, map_pairs = 0;
};
cudaError_t launch_match(sb_point* d_pts1, int n1, const float* d_f1, const sb_point* d_pts2, int n2, const float* d_f2,
                         int nfeatures, MatchScratch& ws, int sm_count, cudaStream_t st);
// pair z: frame d_pairs[2z] against frame d_pairs[2z+1] (d_pairs == nullptr: 2z, 2z+1) of one detect batch, counts on the device
cudaError_t launch_match_batch(sb_point* d_pts, long long pts_stride, const int* d_counts, const float* d_desc, long long desc_stride,
                               int npairs, const int* d_pairs, int bound, int nfeatures, MatchScratch& ws, int sm_count, cudaStream_t st);
void free_match_scratch(MatchScratch& ws);
cudaError_t launch_match_filter(const sb_point* d_pts1, int n1, const sb_point* d_pts2, int n2, float max_ambiguity, int flags,
                                sb_pair* d_pairs, int cap, int* d_count, cudaStream_t st);

}  // namespace sb
