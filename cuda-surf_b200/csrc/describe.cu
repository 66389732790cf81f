// describe.cu -- Haar-wavelet orientation assignment and the 4x4x{4,8} SURF descriptor with fused
// L2 normalisation, ONE WARP PER KEYPOINT.
//
// Replaces cuDescribe (surfd.cu:3251-3325) and its kernels:
//   assignOrientationApprox (surfd.cu:1711-1960), describeURWithoutNormalization (:1566-1615),
//   describeApproxWithoutNormalization (:2391-2444), normalize (:2447-2493).
// The reference sizes a (32x32)-thread grid per keypoint from a global max radius read back to the
// host, scatters every sample into the 64 output floats with float atomicAdd in GLOBAL memory
// (placeInIndex, :1199-1271), then normalises in a third launch. Here:
//   * the keypoint count is read on the device (no host round trip, no cudaMalloc per frame);
//   * a warp walks its keypoint's own (2R+1)^2 sampling lattice, lanes along the lattice row so
//     the integral-image gathers of neighbouring lanes share cache lines;
//   * each lane accumulates into a private copy of the descriptor in shared memory laid out
//     [element][lane] (bank == lane: conflict-free, no atomics, deterministic), reduced across
//     lanes with a rotated read, squared-summed with shuffles and normalised in registers;
//   * descriptors leave the SM once, as 128-byte coalesced stores.
// Arithmetic (rounding modes, FMA placement, IEEE division, __sinf/__cosf) follows the reference
// so that descriptors agree to float round-off.
#include "common.cuh"

namespace sb {

constexpr int kWarpsPerCta = 4;
constexpr int kRotRows = 128;  // lattice rows of the rotated descriptor: side = 2R+1 <= 63 (W = 4) ... 101 (W = 1), since sc/step < 3
constexpr float kR255 = 0.003921568627f;
constexpr float kWindow = 1.0471975511965976f;     // surfd.h:12
constexpr float kSepAngle = 0.08726646259971647f;  // surfd.h:13
constexpr float kHalfPi = 1.5707963267948966f;     // cuda_utils.h:8 (float)
constexpr double kPi = 3.14159265358979323846;     // M_PI (double in the reference)
constexpr int kORadius = 9;                        // surfd.h:15
constexpr int kOSide = 2 * kORadius + 1;           // 19
constexpr int kOSamples = kOSide * kOSide;         // 361
constexpr int kOValid = 253;                       // lattice points with x1^2 + y1^2 < 82 (surfd.cu:1760), walked densely
constexpr int kPasz = kNBin + 2 * kHwn;            // 84

// ---------------------------------------------------------------------------------- orientation

// dFastAtan2, surfd.cu:114-126 (H_PI float, M_PI double).
__device__ __forceinline__ float fast_atan2(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float a = __fdiv_rn(fminf(ax, ay), fmaxf(ax, ay));
    const float s = __fmul_rn(a, a);
    float r = __fmaf_rn(__fmaf_rn(__fmaf_rn(-0.0464964749f, s, 0.15931422f), s, -0.327622764f), __fmul_rn(s, a), a);
    r = (ay > ax) ? __fsub_rn(kHalfPi, r) : r;
    r = (x < 0.f) ? (float)(kPi - (double)r) : r;
    r = (y < 0.f) ? -r : r;
    return r;
}

struct OrientSmem {
    float ang[kOSamples];
    float ps[kOSamples];
    short hid[kOSamples + 1];
    int hist[kNBin];
    int start[kNBin], fill[kNBin];   // run of each bin in `order`
    short order[kOSamples + 3];      // sample indices, stably sorted by bin
    float avg[kNBin];
    float psum[kNBin];
    float pas[kPasz];
    float ws[kNBin];
    float was[kNBin];
};

// grid (ctas, nframes), 4 warps per CTA, warp per keypoint. Deterministic: samples are binned
// in lattice scan order (the reference's shared-memory atomics make its sums order-dependent).
__global__ void __launch_bounds__(kWarpsPerCta * 32)
orient_kernel(const __grid_constant__ PipeP P, const int* __restrict__ Ibase, sb_point* __restrict__ points,
              long long pts_stride, const int* __restrict__ counts, int fixed_count, int* __restrict__ work) {
    pdl_wait();
    __shared__ OrientSmem sm[kWarpsPerCta];
    __shared__ float s_lut1[83];
    __shared__ short s_tab[kOValid + 3];  // lattice index qi of the d-th point inside the circle, ascending
    const int f = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int t = threadIdx.x; t < 83; t += blockDim.x) s_lut1[t] = P.lut1[t];
    if (warp == 0) {
        int nv = 0;
        for (int q0 = 0; q0 < kOSamples; q0 += 32) {
            const int qi = q0 + lane;
            const int y1 = qi / kOSide - kORadius, x1 = qi % kOSide - kORadius;
            const bool in = qi < kOSamples && y1 * y1 + x1 * x1 < 82;
            const unsigned m = __ballot_sync(0xffffffffu, in);
            if (in) s_tab[nv + __popc(m & ((1u << lane) - 1u))] = (short)qi;
            nv += __popc(m);
        }
    }
    __syncthreads();
    const int n = fixed_count >= 0 ? fixed_count : min(counts[f], P.max_pts);
    const int* I = Ibase + (size_t)f * P.istride + P.ip;
    sb_point* pts = points + (size_t)f * pts_stride;
    OrientSmem& S = sm[warp];

    for (int pi = blockIdx.x * kWarpsPerCta + warp; pi < n;) {  // first keypoint fixed, the next ones from the frame's counter
        float x = pts[pi].x, y = pts[pi].y, scale = pts[pi].scale;
        if (P.doubled) { x = __fadd_rn(x, x); y = __fadd_rn(y, y); scale = __fadd_rn(scale, scale); }  // surfd.cu:1734-1739
        const int hs = __float2int_rz(__fmaf_rn(2.f, scale, 1.6f));
        const int st = __float2int_rz(__fadd_rn(scale, 0.8f));
        const int ixc = __float2int_rn(x), iyc = __float2int_rn(y);
        // phase A: Haar responses on the 19x19 lattice
        // (the 253 points inside the circle only, in lattice order: every array below is indexed by that dense position)
        for (int di = lane; di < kOValid; di += 32) {
            const int qi = s_tab[di];
            const int y1 = qi / kOSide - kORadius, x1 = qi % kOSide - kORadius;
            const int xx = ixc + x1 * st, yy = iyc + y1 * st;
            short hid = -1;
            float angle = 0.f, psum = 0.f;
            const int distsq = y1 * y1 + x1 * x1;
            if (yy + hs + 2 < P.ih && yy - hs > -1 && xx + hs + 2 < P.iw && xx - hs > -1) {
                int hx, hy;
                haar_xy(I, P.ip, xx, yy, hs, hx, hy);
                const float dx = __fmul_rn(__int2float_rn(hx), kR255);
                const float dy = __fmul_rn(__int2float_rn(hy), kR255);
                const float mag = __fsqrt_rn(__fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
                if (mag > 0.f) {
                    angle = fast_atan2(dy, dx);
                    hid = (short)(__float2int_rz((float)(((double)angle + kPi) / (double)kSepAngle)) % kNBin);
                    psum = __fmul_rn(s_lut1[distsq], mag);
                }
            }
            S.hid[di] = hid; S.ang[di] = angle; S.ps[di] = psum;
        }
        __syncwarp();
        // phase B: per-bin sums, each bin summed in lattice scan order (the order of the CPU restatement; the
        // reference's shared-memory atomics make its own sums order-dependent). A stable counting sort by bin -- groups
        // of equal bins inside a 32-sample batch found with match.any, their ranks by popc -- lets every lane then walk
        // only the samples of its own bins (3.5 on average) instead of testing all 361 against each bin.
        for (int b = lane; b < kNBin; b += 32) S.hist[b] = 0;
        __syncwarp();
        for (int q0 = 0; q0 < kOValid; q0 += 32) {  // B1: counts
            const int qi = q0 + lane;
            const int hid = qi < kOValid ? (int)S.hid[qi] : -1;
            const unsigned grp = __match_any_sync(0xffffffffu, hid >= 0 ? hid : 128 + lane);
            if (hid >= 0 && (grp & ((1u << lane) - 1u)) == 0) S.hist[hid] += __popc(grp);  // group leader; bins are distinct
            __syncwarp();
        }
        {   // exclusive prefix over the 72 bins -> start of every bin's run in `order`
            const int c0 = S.hist[lane], c1 = S.hist[lane + 32], c2 = lane + 64 < kNBin ? S.hist[lane + 64] : 0;
            int s0 = c0, s1 = c1, s2 = c2;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t0 = __shfl_up_sync(0xffffffffu, s0, d), t1 = __shfl_up_sync(0xffffffffu, s1, d);
                const int t2 = __shfl_up_sync(0xffffffffu, s2, d);
                if (lane >= d) { s0 += t0; s1 += t1; s2 += t2; }
            }
            const int tot0 = __shfl_sync(0xffffffffu, s0, 31), tot1 = __shfl_sync(0xffffffffu, s1, 31);
            S.start[lane] = s0 - c0; S.fill[lane] = s0 - c0;
            S.start[lane + 32] = tot0 + s1 - c1; S.fill[lane + 32] = tot0 + s1 - c1;
            if (lane + 64 < kNBin) { S.start[lane + 64] = tot0 + tot1 + s2 - c2; S.fill[lane + 64] = tot0 + tot1 + s2 - c2; }
        }
        __syncwarp();
        for (int q0 = 0; q0 < kOValid; q0 += 32) {  // B2: stable scatter of the sample indices
            const int qi = q0 + lane;
            const int hid = qi < kOValid ? (int)S.hid[qi] : -1;
            const unsigned grp = __match_any_sync(0xffffffffu, hid >= 0 ? hid : 128 + lane);
            const int rank = __popc(grp & ((1u << lane) - 1u));
            int base = 0;
            if (hid >= 0 && rank == 0) { base = S.fill[hid]; S.fill[hid] = base + __popc(grp); }
            base = __shfl_sync(0xffffffffu, base, __ffs(grp) - 1);
            if (hid >= 0) S.order[base + rank] = (short)qi;
            __syncwarp();
        }
        for (int b = lane; b < kNBin; b += 32) {  // B3: lane owns bins lane, lane+32, lane+64
            const int cnt = S.hist[b], k0 = S.start[b];
            float A = 0.f, Psum = 0.f, Q = 0.f, Qw = 0.f;
            for (int k = 0; k < cnt; k++) {
                const int qi = S.order[k0 + k];
                const float angle = S.ang[qi], ps = S.ps[qi];
                A = __fadd_rn(A, angle);
                Psum = __fadd_rn(Psum, ps);
                Q = __fadd_rn(Q, __fmul_rn(angle, ps));
                if (b < kHwn) Qw = __fadd_rn(Qw, (float)(((double)angle + 2 * kPi) * (double)ps));
                else if (b + kHwn >= kNBin) Qw = __fadd_rn(Qw, (float)(((double)angle - 2 * kPi) * (double)ps));
            }
            S.avg[b] = cnt > 0 ? __fdiv_rn(A, __int2float_rn(cnt)) : P.bins[b];
            S.psum[b] = Psum;
            S.pas[b + kHwn] = Q;
            if (b < kHwn) S.pas[b + kHwn + kNBin] = Qw;
            else if (b + kHwn >= kNBin) S.pas[b + kHwn - kNBin] = Qw;
        }
        __syncwarp();
        // phase C: 60-degree sliding window (+-5 full bins, fractional edge bins), surfd.cu:1848-1908
        for (int i = lane; i < kNBin; i += 32) {
            float ws = 0.f, was = 0.f;
            const float avg_i = S.avg[i];
            for (int j = -kHwn; j <= kHwn; j++) {
                int k = i + j;
                if (j == -kHwn) {
                    float res;
                    if (k < 0) {
                        k += kNBin;
                        const int k1 = (k + 1) % kNBin;
                        const float b1 = P.bins[k1];
                        res = (float)((double)__fsub_rn(__fadd_rn(b1, kWindow / 2), avg_i) - (b1 < 0.f ? 0.0 : 2 * kPi));
                    } else {
                        res = __fsub_rn(__fadd_rn(P.bins[k + 1], kWindow / 2), avg_i);
                    }
                    const float er = __fdiv_rn(res, kSepAngle);
                    ws = __fadd_rn(ws, __fmul_rn(er, S.psum[k]));
                    was = __fadd_rn(was, __fmul_rn(er, S.pas[i]));
                } else if (j == kHwn) {
                    float res;
                    if (k >= kNBin) {
                        k -= kNBin;
                        res = (float)((double)__fadd_rn(avg_i, kWindow / 2) - 2 * kPi - (double)P.bins[k]);
                    } else {
                        res = __fsub_rn(__fadd_rn(avg_i, kWindow / 2), P.bins[k]);
                    }
                    const float er = __fdiv_rn(res, kSepAngle);
                    ws = __fadd_rn(ws, __fmul_rn(er, S.psum[k]));
                    was = __fadd_rn(was, __fmul_rn(er, S.pas[i + 2 * kHwn]));
                } else {
                    was = __fadd_rn(was, S.pas[k + kHwn]);
                    if (k < 0) k += kNBin; else if (k >= kNBin) k -= kNBin;
                    ws = __fadd_rn(ws, S.psum[k]);
                }
            }
            S.ws[i] = ws; S.was[i] = was;
        }
        __syncwarp();
        // phase D: the reference's tournament arg-max (64-wide tree, 8-wide tree at 64, strict <). The pairs of one tree
        // level are disjoint, so a level is one step of the warp (as one lane's loop the 71 compare/copy steps were a
        // fifth of the kernel's instructions and a serial chain of shared-memory round trips).
        {
            int residual = kNBin, offset = 0;
            while (residual > 0) {
                int zn = 1;
                while (zn * 2 <= residual) zn *= 2;
                for (int stride = zn / 2; stride > 0; stride >>= 1) {
                    for (int t = lane; t < stride; t += 32) {
                        const int id1 = t + offset, id2 = id1 + stride;
                        if (S.ws[id1] < S.ws[id2]) { S.ws[id1] = S.ws[id2]; S.was[id1] = S.was[id2]; }
                    }
                    __syncwarp();
                }
                if (lane == 0 && S.ws[0] < S.ws[offset]) { S.ws[0] = S.ws[offset]; S.was[0] = S.was[offset]; }
                __syncwarp();
                residual -= zn;
                offset += zn;
            }
            if (lane == 0) pts[pi].ori = __fdiv_rn(S.was[0], S.ws[0]);
        }
        __syncwarp();
        int nxt = 0;
        if (lane == 0) nxt = gridDim.x * kWarpsPerCta + atomicAdd(work + f, 1);
        pi = __shfl_sync(0xffffffffu, nxt, 0);
    }
}

// ---------------------------------------------------------------------------------- descriptor

// One sample's bilinear spread into the 2x2 nearest cells (placeInIndex, surfd.cu:1199-1271), into this lane's private
// descriptor copy h[element*32 + lane]. Branch-free: a cell outside the WxW grid gets weight 0 and is redirected to a
// per-lane dummy word (row `dummy` of h), so the 8 addresses never alias a live one and the 8 loads are issued together
// before the 8 stores (with a branch per quadrant the updates were 8 dependent shared-memory round trips).
// hs = shared-memory address of this lane's column of the private copy (h + lane), row stride 128 bytes: every update is
// one 32-bit address (cell * O + bin) * 128 + hs -- as generic pointers the eight addresses cost two LEA each on top of the
// index arithmetic.
__device__ __forceinline__ void place(uint32_t hs, int W, int O, int dummy, float mag1, int ori1, float mag2, int ori2, float rx,
                                      float cx) {
    const int ri = __float2int_rz(rx >= 0.f ? rx : __fsub_rn(rx, 1.f));
    const int ci = __float2int_rz(cx >= 0.f ? cx : __fsub_rn(cx, 1.f));
    const float rfrac = __fsub_rn(rx, __int2float_rn(ri));
    const float cfrac = __fsub_rn(cx, __int2float_rn(ci));
    const bool r0ok = ri >= 0, r1ok = ri + 1 < W, c0ok = ci >= 0, c1ok = ci + 1 < W;
    const float rw0 = __fsub_rn(1.f, rfrac), rw1 = rfrac, cw0 = __fsub_rn(1.f, cfrac), cw1 = cfrac;
    const int e00i = (ri * W + ci) * O;  // element index of cell (ri, ci), bin 0
    const int e00 = (r0ok && c0ok) ? e00i : dummy;
    const int e01 = (r0ok && c1ok) ? e00i + O : dummy;
    const int e10 = (r1ok && c0ok) ? e00i + W * O : dummy;
    const int e11 = (r1ok && c1ok) ? e00i + W * O + O : dummy;
    const uint32_t a1 = hs + 128u * (uint32_t)ori1, a2 = hs + 128u * (uint32_t)ori2;
    const uint32_t p[8] = {a1 + 128u * e00, a2 + 128u * e00, a1 + 128u * e01, a2 + 128u * e01,
                           a1 + 128u * e10, a2 + 128u * e10, a1 + 128u * e11, a2 + 128u * e11};
    const float a0 = __fmul_rn(mag1, rw0), b0 = __fmul_rn(mag2, rw0), a1f = __fmul_rn(mag1, rw1), b1 = __fmul_rn(mag2, rw1);
    const float add[8] = {__fmul_rn(a0, cw0), __fmul_rn(b0, cw0), __fmul_rn(a0, cw1), __fmul_rn(b0, cw1),
                          __fmul_rn(a1f, cw0), __fmul_rn(b1, cw0), __fmul_rn(a1f, cw1), __fmul_rn(b1, cw1)};
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[k]) : "r"(p[k]));
#pragma unroll
    for (int k = 0; k < 8; k++) asm volatile("st.shared.f32 [%0], %1;" ::"r"(p[k]), "f"(__fadd_rn(v[k], add[k])) : "memory");
}

// Rotated descriptor (describeApproxWithoutNormalization + addSample, surfd.cu:2391-2444, 1984-2015): the sampling lattice
// stays axis-aligned in the image, the window coordinates (rpos, cpos) are rotated by the keypoint's orientation, so rows
// and columns do not separate as in the upright kernel; a warp walks the (2R+1)^2 lattice 32 samples at a time.
// grid (ctas, nframes), 4 warps per CTA, dynamic smem = 4 * (160 ints + (NF+8)*32 floats) + 40 floats.
__global__ void __launch_bounds__(kWarpsPerCta * 32)
describe_rotated_kernel(const __grid_constant__ PipeP P, const int* __restrict__ Ibase, const sb_point* __restrict__ points,
                        long long pts_stride, const int* __restrict__ counts, int fixed_count, float* __restrict__ desc,
                        long long desc_stride, int* __restrict__ work) {
    pdl_wait();
    extern __shared__ __align__(16) float smem[];
    const int NF = P.nfeatures, W = P.desc_wsz, O = P.orient_size;
    const int f = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // layout: [4 warps][2 * kRotRows + 32] row tables (first column, dense start) | [4 warps][(NF+8)*32] private descriptor
    // copies | lut2[40]
    int* rjlo = reinterpret_cast<int*>(smem) + warp * (2 * kRotRows + 32);
    int* rstart = rjlo + kRotRows;  // kRotRows + 1 entries
    float* hbase = smem + kWarpsPerCta * (2 * kRotRows + 32);
    const int hrows = NF + 8;  // + the dummy rows of place()
    float* s_lut2 = hbase + kWarpsPerCta * hrows * 32;
    for (int t = threadIdx.x; t < 40; t += blockDim.x) s_lut2[t] = P.lut2[t];
    __syncthreads();
    float* h = hbase + warp * hrows * 32;
    const uint32_t hs = (uint32_t)__cvta_generic_to_shared(h + lane);
    const int n = fixed_count >= 0 ? fixed_count : min(counts[f], P.max_pts);
    const int* I = Ibase + (size_t)f * P.istride + P.ip;
    const sb_point* pts = points + (size_t)f * pts_stride;
    float* dout = desc + (size_t)f * desc_stride;
    const float fW = __int2float_rn(W);

    for (int pi = blockIdx.x * kWarpsPerCta + warp; pi < n;) {  // first keypoint fixed, the next ones from the frame's counter
        for (int e = 0; e < hrows; e++) h[e * 32 + lane] = 0.f;
        float x = pts[pi].x, y = pts[pi].y;
        if (P.doubled) { x = __fadd_rn(x, x); y = __fadd_rn(y, y); }  // surfd.cu:2406-2411
        const float sc = __fmul_rn(P.doubled ? 3.3f : 1.65f, pts[pi].scale);
        const int step = max(__float2int_rn(__fmul_rn(sc, 0.5f)), 1);
        const int ixc = __float2int_rn(x), iyc = __float2int_rn(y);
        const float fx = __fsub_rn(x, __int2float_rn(ixc)), fy = __fsub_rn(y, __int2float_rn(iyc));
        const float spacing = __fmul_rn(sc, __int2float_rn(P.mag_factor));
        const int S = __float2int_rz(sc);
        const float wofs = __fmaf_rn(fW, 0.5f, -0.5f);
        const float fstep = __int2float_rn(step);
        const float ori = pts[pi].ori;
        const float sine = __sinf(ori), cose = __cosf(ori);
        const float fracc = __fmaf_rn(-sine, fy, __fmul_rn(cose, fx));
        const float fracr = __fmaf_rn(cose, fy, __fmul_rn(sine, fx));
        const int R = __float2int_rn(__fdiv_rn(__fmul_rn(__fmul_rn(__fmul_rn(1.4f, spacing), __int2float_rn(W + 1)), 0.5f), fstep));
        const int side = 2 * R + 1;
        // |rpos| < (W+1)/2 is the window test -1 < rpos + wofs < W; about half of the lattice (the corners of the bounding
        // square of the rotated window) fails it by a wide margin and is rejected on the numerators, before the two IEEE
        // divisions. The margin keeps the exact test below the only one that decides.
        const float far_lim = __fmul_rn(__fmaf_rn(fW, 0.5f, 0.51f), spacing);
        // Enumeration of the lattice. The rotated window covers about half of its bounding (2R+1)^2 lattice, so walking
        // all of it left half of every warp idle (and the rejection tests alone were half of the instructions). Here each
        // lattice ROW gets the interval of columns that can pass -- |nr|, |nc| <= far_lim are linear in j, plus the image
        // border in c, widened by one column on both sides against rounding -- a warp scan of the interval lengths makes a
        // dense index, and the warp walks that: a superset of the valid points, in lattice order, ~95 % of the lanes
        // passing the exact tests (unchanged) below.
        if (side > kRotRows) __trap();  // cannot happen: R = rn(1.4 * 6 sc (W+1)/W / step) with sc / step < 3
        int total = 0;
        {
            const float A = __fmul_rn(fstep, sine), B = __fmul_rn(fstep, cose);
            int cnt[kRotRows / 32];
#pragma unroll
            for (int u = 0; u < kRotRows / 32; u++) {
                const int ii = lane + 32 * u;
                int lo = 0, hi = -1;
                if (ii < side) {
                    const float fi = __int2float_rn(ii - R);
                    const float cr = B * fi - fracr, cc = -A * fi - fracc;  // nr = cr + A j, nc = cc + B j
                    float flo = (float)-R, fhi = (float)R;
                    if (fabsf(A) > 1e-6f) {
                        const float t0 = (-far_lim - cr) / A, t1 = (far_lim - cr) / A;
                        flo = fmaxf(flo, fminf(t0, t1)); fhi = fminf(fhi, fmaxf(t0, t1));
                    } else if (fabsf(cr) > far_lim) { fhi = flo - 4.f; }
                    if (fabsf(B) > 1e-6f) {
                        const float t0 = (-far_lim - cc) / B, t1 = (far_lim - cc) / B;
                        flo = fmaxf(flo, fminf(t0, t1)); fhi = fminf(fhi, fmaxf(t0, t1));
                    } else if (fabsf(cc) > far_lim) { fhi = flo - 4.f; }
                    // image border: 1 + S <= ixc + j*step < iw - 1 - S
                    flo = fmaxf(flo, (float)(1 + S - ixc) / fstep);
                    fhi = fminf(fhi, (float)(P.iw - 2 - S - ixc) / fstep);
                    const int r = iyc + (ii - R) * step;
                    if (!(r >= 1 + S && r < P.ih - 1 - S)) fhi = flo - 4.f;
                    lo = max(-R, (int)floorf(flo) - 1);
                    hi = min(R, (int)ceilf(fhi) + 1);
                }
                cnt[u] = max(0, hi - lo + 1);
                rjlo[ii] = lo;
            }
            // exclusive scan over the rows (row ii = lane + 32 u)
            int run = 0;
#pragma unroll
            for (int u = 0; u < kRotRows / 32; u++) {
                int incl = cnt[u];
#pragma unroll
                for (int o2 = 1; o2 < 32; o2 <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o2);
                    if (lane >= o2) incl += t;
                }
                rstart[lane + 32 * u] = run + incl - cnt[u];
                run += __shfl_sync(0xffffffffu, incl, 31);
            }
            total = run;
            if (lane == 0) rstart[kRotRows] = total;
        }
        __syncwarp();
        int row = 0;
        for (int d0 = 0; d0 < total; d0 += 32) {
            const int d = d0 + lane;
            if (d < total) {
                while (d >= rstart[row + 1]) row++;
                const int i = row - R, j = rjlo[row] + (d - rstart[row]);
                const float fi = __int2float_rn(i), fj = __int2float_rn(j);
                const float nr = __fmaf_rn(fstep, __fmaf_rn(cose, fi, __fmul_rn(sine, fj)), -fracr);
                const float nc = __fmaf_rn(fstep, __fmaf_rn(-sine, fi, __fmul_rn(cose, fj)), -fracc);
                if (!(fabsf(nr) > far_lim || fabsf(nc) > far_lim)) {
                    const float rpos = __fdiv_rn(nr, spacing), cpos = __fdiv_rn(nc, spacing);
                    const float rx = __fadd_rn(rpos, wofs), cx = __fadd_rn(cpos, wofs);
                    const int r = iyc + i * step, c = ixc + j * step;
                    if (rx > -1.f && rx < fW && cx > -1.f && cx < fW && r >= 1 + S && r < P.ih - 1 - S && c >= 1 + S &&
                        c < P.iw - 1 - S) {
                        const float weight = s_lut2[__float2int_rz(__fmaf_rn(rpos, rpos, __fmul_rn(cpos, cpos)))];
                        int hx, hy;
                        haar_xy(I, P.ip, c, r, S, hx, hy);
                        const float a = __fmul_rn(__fmul_rn(weight, __int2float_rn(hx)), kR255);
                        const float b = __fmul_rn(__fmul_rn(weight, __int2float_rn(hy)), kR255);
                        const float dx = __fmaf_rn(cose, a, __fmul_rn(sine, b));
                        const float dy = __fmaf_rn(sine, a, -__fmul_rn(cose, b));
                        if (O == 4) {
                            place(hs, W, O, NF, dx, (dx < 0.f ? 0 : 1), dy, (dy < 0.f ? 2 : 3), rx, cx);
                        } else {
                            place(hs, W, O, NF, dx, (dy < 0.f ? 0 : 1), fabsf(dx), (dy < 0.f ? 2 : 3), rx, cx);
                            place(hs, W, O, NF, dy, (dx < 0.f ? 4 : 5), fabsf(dy), (dx < 0.f ? 6 : 7), rx, cx);
                        }
                    }
                }
            }
        }
        __syncwarp();
        // reduce the 32 private copies (rotated read: bank == (lane + k) % 32), normalise, store
        float v[4];
        float sq = 0.f;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int e = lane + 32 * u;
            float acc = 0.f;
            if (e < NF) {
                for (int k = 0; k < 32; k++) acc += h[e * 32 + ((lane + k) & 31)];
            }
            v[u] = acc;
            sq = __fmaf_rn(acc, acc, sq);
        }
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o2);
        const float inv = __fdiv_rn(1.f, __fsqrt_rn(sq));
        float* d = dout + (size_t)pi * NF;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int e = lane + 32 * u;
            if (e < NF) d[e] = __fmul_rn(v[u], inv);
        }
        __syncwarp();
        int nxt = 0;
        if (lane == 0) nxt = gridDim.x * kWarpsPerCta + atomicAdd(work + f, 1);
        pi = __shfl_sync(0xffffffffu, nxt, 0);
    }
}


// ------------------------------------------------------------------ upright descriptor, v2
//
// Upright sampling is separable: rpos depends only on the lattice row, cpos only on the column.
// A lane therefore OWNS a lattice column (its cpos, cell column ci, bilinear weights and the four
// integral-image column offsets are computed once per pass) and sweeps the rows; the per-row
// quantities come from a small per-keypoint table built once by the whole warp. Because the cell
// row ri only grows along the sweep, the contributions to cell rows ri / ri+1 live in registers
// and are flushed (x the column weights) into the lane's private descriptor copy only when ri
// advances: ~10 shared-memory updates per column instead of 8 per sample. The two half-warps take
// even / odd lattice rows, 16 columns per pass, so 23..45-wide lattices keep 70-97 % of the lanes
// busy. 12 integral loads per sample (the reference's two Haar boxes share four corners).
struct __align__(16) RowEntry { float rpos, w1, w0; int rowoff; };  // w1 = bilinear weight of cell row ri+1, w0 = 1 - w1
constexpr int kRowTab = 96;
constexpr int kAhead = 1;  // rows (of one parity) the gathers run ahead of the arithmetic
constexpr int kRowInvalid = -100;  // 'no current cell row' marker of the sweep

// Row stride of the per-warp staging tile T[cell row * O + bin][lane] and of the column-weight table Wc[cell column][lane]:
// 36 floats = 16-byte aligned rows whose banks advance by 4 per row, so the reduction's LDS.128 of 8 different rows
// (one per lane group) tile the 32 banks exactly.
constexpr int kTS = 36;

struct TrueT { static constexpr bool value = true; };
struct FalseT { static constexpr bool value = false; };

// LIST: the keypoints come from the class list the TMA path left (describe_tma.cu); false compiles the list handling away
template <int O, bool LIST>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 5)
describe_upright_kernel(const __grid_constant__ PipeP P, const int* __restrict__ Ibase, const sb_point* __restrict__ points,
                        long long pts_stride, const int* __restrict__ counts, int fixed_count, float* __restrict__ desc,
                        long long desc_stride, int* __restrict__ work, const int* __restrict__ cls_idx, int* __restrict__ cls_cnt,
                        int slot0) {
    pdl_wait();
    extern __shared__ __align__(16) float smem[];
    const int NF = P.nfeatures, W = P.desc_wsz;
    const int f = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // layout: [4 warps][(W*O + W) * kTS] staging tile + column weights | [4 warps][kRowTab] row tables | lut2[40]
    const int stage = (W * O + W) * kTS;
    RowEntry* rowT = reinterpret_cast<RowEntry*>(smem + kWarpsPerCta * stage) + warp * kRowTab;
    float* s_lut2 = smem + kWarpsPerCta * stage + kWarpsPerCta * kRowTab * 4;
    for (int t = threadIdx.x; t < 40; t += blockDim.x) s_lut2[t] = P.lut2[t];
    // the tile starts finite: lanes without a valid column never write it and enter the reduction with weight 0
    for (int t = threadIdx.x; t < kWarpsPerCta * stage; t += blockDim.x) smem[t] = 0.f;
    __syncthreads();
    float* T = smem + warp * stage;
    float* Wc = T + W * O * kTS;
    const unsigned lut_sa = (unsigned)__cvta_generic_to_shared(s_lut2);
    // with a class list (describe_tma.cu) this kernel takes the keypoints the TMA path left: list position -> keypoint
    const int* list = LIST ? cls_idx + ((size_t)(slot0 + f) * 2 + 1) * P.max_pts : nullptr;
    int* wk = LIST ? cls_cnt + (slot0 + f) * 4 + 3 : work + f;  // this frame's work counter
    const int n = LIST ? cls_cnt[(slot0 + f) * 4 + 1] : (fixed_count >= 0 ? fixed_count : min(counts[f], P.max_pts));
    const int ip = P.ip;
    const int* I = Ibase + (size_t)f * P.istride + ip;
    const sb_point* pts = points + (size_t)f * pts_stride;
    float* dout = desc + (size_t)f * desc_stride;
    const float fW = __int2float_rn(W);
    const int half = lane >> 4, jl = lane & 15;

    // a warp's first keypoint is fixed, the following ones come from a per-frame counter (zero on entry): keypoints cost
    // 2 or 3 passes, and with one or two per warp (single frame) a static split left the machine waiting for the unlucky warps
    for (int pk = blockIdx.x * kWarpsPerCta + warp; pk < n;) {
        const int pi = LIST ? list[pk] : pk;
        const float x = pts[pi].x, y = pts[pi].y;
        const KpGeom kg = kp_geom(x, y, pts[pi].scale, W, P.mag_factor, P.doubled);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};  // descriptor elements lane, lane+32, ... (un-normalised)
        const int step = kg.step, ixc = kg.ixc, iyc = kg.iyc, S = kg.S, R = kg.R;
        const float fx = kg.fx, fy = kg.fy, spacing = kg.spacing;
        const float wofs = __fmaf_rn(fW, 0.5f, -0.5f);
        const int side = min(2 * R + 1, kRowTab);
        // per-row table (same IEEE operations as the reference's per-sample arithmetic). Valid rows
        // (inside the descriptor window and the image) form one contiguous range [row_lo, row_hi).
        int row_lo = side, row_hi = 0;
        // segcnt[s] = number of valid rows with ri <= s-1: the rows of cell row ri = s-1 are [row_lo + segcnt[s-1],
        // row_lo + segcnt[s]) (ri only grows with the row), so the sweep flushes at segment ends, not per sample
        int segcnt[5] = {0, 0, 0, 0, 0};
        for (int i0 = 0; i0 < side; i0 += 32) {
            const int ii = i0 + lane;
            const int i = ii - R;
            const float rpos = __fdiv_rn(__fsub_rn(__int2float_rn(step * i), fy), spacing);
            const float rx = __fadd_rn(rpos, wofs);
            const int r = iyc + i * step;
            const bool ok = ii < side && rx > -1.f && rx < fW && r >= 1 + S && r < P.ih - 1 - S;
            const int ri = __float2int_rz(rx >= 0.f ? rx : __fsub_rn(rx, 1.f));
            if (ii < side) {
                RowEntry t;
                t.rpos = rpos; t.w1 = __fsub_rn(rx, __int2float_rn(ri)); t.w0 = __fsub_rn(1.f, t.w1); t.rowoff = r * ip;
                rowT[ii] = t;
            }
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (m) {
                row_lo = min(row_lo, i0 + __ffs(m) - 1);
                row_hi = max(row_hi, i0 + 32 - __clz(m));
            }
#pragma unroll
            for (int sg = 0; sg < 5; sg++) segcnt[sg] += __popc(__ballot_sync(0xffffffffu, ok && ri <= sg - 1));
        }
        __syncwarp();
        const int sip = S * ip, sip1 = sip + ip;
        for (int p0 = 0; p0 < side; p0 += 16) {
            const int jj = p0 + jl;
            const int j = jj - R;
            const float cpos = __fdiv_rn(__fsub_rn(__int2float_rn(step * j), fx), spacing);
            const float cx = __fadd_rn(cpos, wofs);
            const int c = ixc + j * step;
            const bool colok = jj < side && cx > -1.f && cx < fW && c >= 1 + S && c < P.iw - 1 - S;
            const int ci = __float2int_rz(cx >= 0.f ? cx : __fsub_rn(cx, 1.f));
            const float cfrac = __fsub_rn(cx, __int2float_rn(ci)), cfrac1 = __fsub_rn(1.f, cfrac);
            if (colok) {
                const float cpos2 = __fmul_rn(cpos, cpos);
                for (int t = 0; t < W * O; t++) T[t * kTS + lane] = 0.f;  // cell rows this lane's sweep never reaches
                // Column base pointers, made opaque so every gather is ONE IMAD.WIDE (row offset * 4 + base)
                // instead of a 64-bit add chain per load.
                const int* pA = I + (c - S);
                const int* pB = I + c;
                const int* pD = I + (c + S + 1);
                asm volatile("" : "+l"(pA), "+l"(pB), "+l"(pD));
                // Register accumulators for cell rows ri (lo) and ri+1 (hi): signed sum and sum of magnitudes of dx and dy;
                // the reference's split by sign is (S -/+ A)/2 at flush time.
                float lo[4] = {0.f, 0.f, 0.f, 0.f}, hi[4] = {0.f, 0.f, 0.f, 0.f};
                float xlo[4] = {0.f, 0.f, 0.f, 0.f}, xhi[4] = {0.f, 0.f, 0.f, 0.f};  // SURF-128: the sums restricted to dy<0 / dx<0
                auto flush = [&](int k, const float (&sv)[4], const float (&xv)[4]) {
                    if (k < 0 || k >= W) return;
                    float v[O];
                    if (O == 4) {
                        // staged as signed sum / sum of magnitudes; the split by sign (addUprightSample, surfd.cu:1308:
                        // bins dx<0, dx>=0, dy<0, dy>=0 = (S -/+ A)/2) is linear and applied once, after the reduction
                        v[0] = sv[0]; v[1] = sv[1]; v[2] = sv[2]; v[3] = sv[3];
                    } else {
                        // SURF-128 (surfd.cu:1312-1313): dx,|dx| split by sign of dy; dy,|dy| split by sign of dx
                        v[0] = xv[0]; v[1] = sv[0] - xv[0]; v[2] = xv[1]; v[3] = sv[1] - xv[1];
                        v[4] = xv[2]; v[5] = sv[2] - xv[2]; v[6] = xv[3]; v[7] = sv[3] - xv[3];
                    }
                    // staged UNWEIGHTED: the split over the cell columns ci, ci+1 is applied by the reduction below
                    // (read-modify-write of two private descriptor cells per value was 30-40 % of a pass)
                    float* t = T + (k * O) * kTS + lane;
#pragma unroll
                    for (int b = 0; b < O; b++) t[b * kTS] = v[b];
                };
                // The 12 corner values of a sample: rows r-S (m), r (z), r+1 (u), r+S+1 (q); columns c-S (A), c (B),
                // c+1 (C), c+S+1 (D); m and q need all four columns, z and u only A and D.
                // Because step = rn(sc/2) and S = rz(sc), S = 2*step - e with e in {0,1}, and this lane visits
                // lattice rows i, i+2, ... (pixel rows 2*step apart). Hence the row r-S of this iteration is the
                // row X = r+e of the previous one, and the row Y = r+1-e of this iteration is the previous row
                // r+S+1: 6 of the 12 values are carried in registers and an iteration gathers 8 (rows X and q),
                //   haar_x = (qD + mB - mD - qB) - (qC + mA - mC - qA)
                //   haar_y = (XD - XA) + (YD - YA) - (mD - mA) - (qD - qA)      (symmetric in z/u, so X/Y need no
                // case split).
                // Software pipeline without register moves: three register sets (8 corners + the row's rpos, w0, w1)
                // rotate through the roles previous / current / next, so the loop body is unrolled three times; the
                // gathers of the NEXT row are issued before the current row is consumed.
                const int e = 2 * step - S;
                const bool carry = (e == 0 || e == 1);
                const int offX = carry ? e * ip : 0, offY = carry ? ip - offX : ip;
                struct RowSet { int g[8]; float rpos, w0, w1; int rowoff; };
                auto gather8 = [&](int ii, RowSet& r) {
                    const RowEntry t = rowT[ii];
                    r.rpos = t.rpos; r.w0 = t.w0; r.w1 = t.w1; r.rowoff = t.rowoff;
                    const int ox = t.rowoff + offX, op = t.rowoff + sip1;
                    r.g[0] = __ldg(pA + ox); r.g[1] = __ldg(pB + ox); r.g[2] = __ldg(pB + ox + 1); r.g[3] = __ldg(pD + ox);
                    r.g[4] = __ldg(pA + op); r.g[5] = __ldg(pB + op); r.g[6] = __ldg(pB + op + 1); r.g[7] = __ldg(pD + op);
                };
                int ii = row_lo + half;
                int seg = 0, segend = row_lo + segcnt[0];
                // one sample: prev = row two lattice steps up (its X row is this row's m, its q row this row's Y)
                // `carry` is uniform over the warp (one keypoint), so it selects one of two instantiations of the sweep:
                // as a run-time predicate the re-read of the generic case cost 17 issue slots per sample even when off.
                auto sample = [&](auto carry_tag, const RowSet& prev, const RowSet& cur, RowSet& next) {
                    constexpr bool kCarry = decltype(carry_tag)::value;
                    if (ii + 2 * kAhead < row_hi) gather8(ii + 2 * kAhead, next);
                    while (ii >= segend) {  // the rows of cell row seg-1 are done: flush it, the upper half moves down
                        if (seg >= 1) flush(seg - 1, lo, xlo);
#pragma unroll
                        for (int b = 0; b < 4; b++) { lo[b] = hi[b]; hi[b] = 0.f; xlo[b] = xhi[b]; xhi[b] = 0.f; }
                        seg++;
                        // (selects, not an indexed array: segcnt stays in registers)
                        const int cnt = seg == 1 ? segcnt[1] : seg == 2 ? segcnt[2] : seg == 3 ? segcnt[3] : segcnt[4];
                        segend = seg <= W ? row_lo + cnt : (1 << 30);
                    }
                    int m0 = prev.g[0], m1 = prev.g[1], m2 = prev.g[2], m3 = prev.g[3], yA = prev.g[4], yD = prev.g[7];
                    if (!kCarry) {
                        // generic (S, step): rows z = r and u = r+1 were gathered as X and Y, m is re-read
                        const int om = cur.rowoff - sip, oy = cur.rowoff + offY;
                        m0 = __ldg(pA + om); m1 = __ldg(pB + om); m2 = __ldg(pB + om + 1); m3 = __ldg(pD + om);
                        yA = __ldg(pA + oy); yD = __ldg(pD + oy);
                    }
                    float weight;
                    asm("ld.shared.f32 %0, [%1];" : "=f"(weight) : "r"(lut_sa + 4u * (unsigned)__float2int_rz(__fmaf_rn(cur.rpos, cur.rpos, cpos2))));
                    const int wx = (cur.g[7] + m1 - m3 - cur.g[5]) - (cur.g[6] + m0 - m2 - cur.g[4]);
                    const int wy = (cur.g[3] - cur.g[0]) + (yD - yA) - (m3 - m0) - (cur.g[7] - cur.g[4]);
                    const float a = __fmul_rn(__fmul_rn(weight, __int2float_rn(wx)), kR255);
                    const float b = __fmul_rn(__fmul_rn(weight, __int2float_rn(wy)), kR255);
                    const float w0 = cur.w0, w1 = cur.w1;
                    lo[0] = __fmaf_rn(a, w0, lo[0]); lo[1] = __fmaf_rn(fabsf(a), w0, lo[1]);
                    lo[2] = __fmaf_rn(b, w0, lo[2]); lo[3] = __fmaf_rn(fabsf(b), w0, lo[3]);
                    hi[0] = __fmaf_rn(a, w1, hi[0]); hi[1] = __fmaf_rn(fabsf(a), w1, hi[1]);
                    hi[2] = __fmaf_rn(b, w1, hi[2]); hi[3] = __fmaf_rn(fabsf(b), w1, hi[3]);
                    if (O == 8) {
                        // the part of each sum with (dy<0) for the dx-sums, (dx<0) for the dy-sums
                        const float an = (b < 0.f) ? a : 0.f, bn = (a < 0.f) ? b : 0.f;
                        xlo[0] = __fmaf_rn(an, w0, xlo[0]); xlo[1] = __fmaf_rn(fabsf(an), w0, xlo[1]);
                        xlo[2] = __fmaf_rn(bn, w0, xlo[2]); xlo[3] = __fmaf_rn(fabsf(bn), w0, xlo[3]);
                        xhi[0] = __fmaf_rn(an, w1, xhi[0]); xhi[1] = __fmaf_rn(fabsf(an), w1, xhi[1]);
                        xhi[2] = __fmaf_rn(bn, w1, xhi[2]); xhi[3] = __fmaf_rn(fabsf(bn), w1, xhi[3]);
                    }
                    ii += 2;
                };
                if (ii < row_hi) {
                    RowSet r0, r1, r2;
                    gather8(ii, r1);
                    {   // prologue: the m row (r-S) and the Y row (r+1-e) of the first sample, in the layout of a previous set
                        const int om = r1.rowoff - sip, oy = r1.rowoff + offY;
                        r0.g[0] = __ldg(pA + om); r0.g[1] = __ldg(pB + om); r0.g[2] = __ldg(pB + om + 1); r0.g[3] = __ldg(pD + om);
                        r0.g[4] = __ldg(pA + oy); r0.g[7] = __ldg(pD + oy); r0.g[5] = 0; r0.g[6] = 0;
                        r0.rpos = 0.f; r0.w0 = 0.f; r0.w1 = 0.f; r0.rowoff = 0;
                    }
                    if (carry) {
                        while (true) {
                            sample(TrueT{}, r0, r1, r2); if (ii >= row_hi) break;
                            sample(TrueT{}, r1, r2, r0); if (ii >= row_hi) break;
                            sample(TrueT{}, r2, r0, r1); if (ii >= row_hi) break;
                        }
                    } else {
                        while (true) {
                            sample(FalseT{}, r0, r1, r2); if (ii >= row_hi) break;
                            sample(FalseT{}, r1, r2, r0); if (ii >= row_hi) break;
                            sample(FalseT{}, r2, r0, r1); if (ii >= row_hi) break;
                        }
                    }
                    // the last row's cell rows: ri = seg-1 (lo) and ri+1 (hi)
                    if (seg >= 1) flush(seg - 1, lo, xlo);
                    flush(seg, hi, xhi);
                }
            }
            // Column weights of this pass (placeInIndex's split over cell columns ci and ci+1, surfd.cu:1199-1271), then
            // element e = (k*W + cc)*O + o gathers sum_l Wc[cc][l] * T[k*O + o][l] over the 32 lanes' columns.
            for (int cc = 0; cc < W; cc++)
                Wc[cc * kTS + lane] = colok ? ((ci == cc ? cfrac1 : 0.f) + (ci + 1 == cc ? cfrac : 0.f)) : 0.f;
            __syncwarp();
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int e = lane + 32 * u;
                if (e < NF) {
                    const int o = e % O, t = e / O;
                    const int k = (t >= W) + (t >= 2 * W) + (t >= 3 * W), cc = t - k * W;
                    const float4* tr = reinterpret_cast<const float4*>(T + (k * O + o) * kTS);
                    const float4* wr = reinterpret_cast<const float4*>(Wc + cc * kTS);
                    float a = acc[u];
#pragma unroll
                    for (int l4 = 0; l4 < 8; l4++) {
                        const float4 tv = tr[l4], wv = wr[l4];
                        a = __fmaf_rn(wv.x, tv.x, a); a = __fmaf_rn(wv.y, tv.y, a);
                        a = __fmaf_rn(wv.z, tv.z, a); a = __fmaf_rn(wv.w, tv.w, a);
                    }
                    acc[u] = a;
                }
            }
            __syncwarp();  // the next pass overwrites T and Wc
        }
        // split by sign (O == 4: elements o = 0,1 hold S, A of dx and become (S-A)/2, (S+A)/2; o = 2,3 likewise for dy: the
        // partner is the neighbouring lane), normalise, store
        float v[4];
        float sq = 0.f;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            float a = acc[u];
            if (O == 4) {
                const float pr = __shfl_xor_sync(0xffffffffu, a, 1);
                a = (lane & 1) ? 0.5f * (pr + a) : 0.5f * (a - pr);
            }
            v[u] = a;
            sq = __fmaf_rn(a, a, sq);
        }
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o2);
        const float inv = __fdiv_rn(1.f, __fsqrt_rn(sq));
        float* d = dout + (size_t)pi * NF;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int e = lane + 32 * u;
            if (e < NF) d[e] = __fmul_rn(v[u], inv);
        }
        __syncwarp();
        int nxt = 0;
        if (lane == 0) nxt = gridDim.x * kWarpsPerCta + atomicAdd(wk, 1);
        pk = __shfl_sync(0xffffffffu, nxt, 0);
    }
}

cudaError_t launch_describe(const PipeP& P, int nframes, const int* d_integral, sb_point* d_points, long long pts_stride,
                            const int* d_counts, int fixed_count, float* d_desc, long long desc_stride, int sm_count,
                            int* d_work, int* d_work_orient, const DescAux& aux, cudaStream_t st) {
    const int maxn = fixed_count >= 0 ? fixed_count : P.max_pts;
    if (maxn <= 0 || nframes <= 0) return cudaSuccess;
    // upright 64-d: the keypoints with sampling step 2 go through the TMA-staged kernel, the gather kernel below takes the rest
    const bool tma = aux.maps && describe_tma_applies(P);
    if (tma) {
        const cudaError_t e = launch_describe_tma(P, nframes, aux, d_points, pts_stride, d_counts, fixed_count, d_desc, desc_stride, sm_count, st);
        if (e != cudaSuccess) return e;
    }
    const int* cls_idx = tma ? aux.cls_idx : nullptr;
    int* cls_cnt = tma ? aux.cls_cnt : nullptr;
    const int need = (maxn + kWarpsPerCta - 1) / kWarpsPerCta;
    // blockIdx.x runs fastest, so 1-2 CTAs per SM per frame keep only a few frames' integral images
    // live at a time (L2-resident; measured: 1 is best from 32 frames, 2 at 8) while still covering the machine for a single frame
    // (a single frame or a small batch fills the machine instead: 5 resident CTAs per SM)
    int ctas = sm_count * (nframes >= 32 ? 1 : nframes >= 5 ? 2 : (nframes >= 3 ? 3 : 5));
    if (ctas < 1) ctas = 1;
    if (ctas > need) ctas = need;
    const dim3 grid(ctas, nframes), block(kWarpsPerCta * 32);
    // (the orientation pass has its own counters: d_work_orient)
    cudaError_t e = cudaSuccess;
    if (!P.upright) e = launch_dep(orient_kernel, grid, block, 0, st, P, d_integral, d_points, pts_stride, d_counts, fixed_count, d_work_orient);
    if (e != cudaSuccess) return e;
    if (aux.before_last && (e = aux.before_last(aux.before_last_arg)) != cudaSuccess) return e;
    if (P.upright) {
        const size_t smem = ((size_t)kWarpsPerCta * (P.desc_wsz * P.orient_size + P.desc_wsz) * kTS + kWarpsPerCta * kRowTab * 4 + 40) * sizeof(float);
        if (P.orient_size == 4) {
            if (tma) {
                cudaFuncSetAttribute(describe_upright_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                return launch_dep(describe_upright_kernel<4, true>, grid, block, smem, st, P, d_integral, d_points, pts_stride, d_counts, fixed_count, d_desc, desc_stride, d_work, cls_idx, cls_cnt, aux.slot0);
            }
            cudaFuncSetAttribute(describe_upright_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            return launch_dep(describe_upright_kernel<4, false>, grid, block, smem, st, P, d_integral, d_points, pts_stride, d_counts, fixed_count, d_desc, desc_stride, d_work, (const int*)nullptr, (int*)nullptr, 0);
        } else {
            cudaFuncSetAttribute(describe_upright_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            return launch_dep(describe_upright_kernel<8, false>, grid, block, smem, st, P, d_integral, d_points, pts_stride, d_counts, fixed_count, d_desc, desc_stride, d_work, (const int*)nullptr, (int*)nullptr, 0);
        }
    } else {
        const size_t smem = ((size_t)kWarpsPerCta * (2 * kRotRows + 32) + (size_t)kWarpsPerCta * (P.nfeatures + 8) * 32 + 40) * sizeof(float);
        cudaFuncSetAttribute(describe_rotated_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        return launch_dep(describe_rotated_kernel, grid, block, smem, st, P, d_integral, d_points, pts_stride, d_counts, fixed_count, d_desc, desc_stride, d_work);
    }
    return cudaGetLastError();
}

}  // namespace sb
