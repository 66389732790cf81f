// describe_tma.cu -- upright descriptor from TMA-staged integral patches (the most frequent keypoint geometry).
//
// Replaces describeURWithoutNormalization + normalize (surfd.cu:1566-1615, 2447-2493) for keypoints whose sampling
// step is 2 (58 % of the keypoints of a 1080p frame; every third one of them has R > 15 and stays with the gather kernel
// of describe.cu, like the other steps -- the split is made on the device by classify_kernel).
//
// Why: the gather kernel reads 8 scattered 4-byte words per sample through the L1 (10 sectors per request, 68 % of the
// L1 pipe). Here ONE WARP owns a keypoint and the TMA stages the patch of the integral image the keypoint needs:
//   * S = 2*step - e (e in {0,1}), so every corner of every Haar box lies on lattice row r0 + 2k (+1) -- the two ROW
//     PHASES are de-interleaved by the TMA itself (elementStrides[1] = 2: box rows are every second image row); along x
//     the TMA cannot stride (cuda.h: elementStrides[0] is ignored), so columns stay dense, which for step 2 is exactly
//     what is needed: lattice columns c0 + 2j and c0 + 2j + 1 are all columns;
//   * boxes are 16 ints wide with the 64-byte swizzle, so a LANE THAT OWNS A LATTICE ROW reads four consecutive
//     columns of its row with one conflict-free LDS.128 (8 lanes = 8 consecutive rows = 8 different 16-byte slots);
//   * the lane sweeps its row left to right with a register window of 4 chunks x 4 row planes (rows r-S, r, r+1, r+S+1);
//     one chunk step = 4 LDS.128 = two samples, each sample 8 IADD3 for both Haar responses (the gather kernel: 8 LDG +
//     8 IMAD.WIDE + 14 integer operations);
//   * column quantities (cpos^2, bilinear column weights, cell column) come from a per-keypoint table (broadcast
//     LDS.128), row quantities are per-lane constants; the bilinear split over cell columns lives in registers and is
//     staged at cell-column boundaries, the split over cell rows is applied by the 32-lane reduction -- the mirror image
//     of describe_upright_kernel;
//   * the patch of the NEXT keypoint is fetched into the same buffer block by block as the sweep leaves a block
//     (mbarrier per block), so the TMA latency (1 us for 20 KB) is hidden behind the sweep.
// Box start columns must be multiples of 4 ints (an unaligned start raises an illegal-instruction fault, measured with
// tools/tma_probe.cu), so the patch starts at the aligned column left of the lattice and the two possible alignments
// of the lattice inside the chunks are two instantiations of the sweep (x2 for e).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace sb {

namespace {

constexpr float kR255 = 0.003921568627f;
constexpr int kTS = 36;                 // row stride (floats) of the staging tile and the row-weight table, as in describe.cu
constexpr int kFK = 36;                 // rows per (block, phase) slot: side + 4 <= 35
constexpr int kFNB = 5;                 // 16-int column blocks
constexpr int kFSlot = kFK * 64;        // 2304 B = 18 * 128: the 64B-swizzle phase of a row is ((k >> 1) & 3) ^ (phase << 1)
constexpr int kFPatch = 2 * kFNB * kFSlot;
constexpr int kFMaxSide = 31;           // R <= 15
constexpr int kFColTab = 40;            // virtual lattice columns (side + 1 <= 32) + slack
// shared memory of one warp: patch | T[16][kTS] | Wr[4][kTS] | column tables float4 + float [kFColTab] | lut2[40] | 5 mbarriers
constexpr int kFSmem = kFPatch + (16 * kTS + 4 * kTS) * 4 + kFColTab * 20 + 40 * 4 + 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded: a mis-programmed pipeline traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 22); it++) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
    }
    __trap();
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ int4 lds128(uint32_t a) {
    int4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ int elem(const int4& v, int e) { return e == 0 ? v.x : e == 1 ? v.y : e == 2 ? v.z : v.w; }

// which keypoints take this path (uniform over the keypoint): upright 64-d only, step 2, e in {0, 1}, side <= 31
__device__ __forceinline__ bool fast_geom(const KpGeom& g) { return g.step == 2 && (g.e == 0 || g.e == 1) && g.side <= kFMaxSide; }

// patch placement of a keypoint: aligned start column, first TMA row (phase 0), box rows and blocks
struct Patch { int xs, ys, krows, nblk, t0, odd, tsteps; };
__device__ __forceinline__ Patch patch_of(const KpGeom& g) {
    Patch p;
    const int col_first = g.ixc - 2 * g.R;                 // integral column `c` of lattice column jj = 0
    p.t0 = (col_first & 3) >= 2 ? 1 : 0;                    // pairs start one (phantom) column to the left
    const int cstar = col_first - 2 * p.t0;
    p.odd = cstar & 3;                                      // 0 or 1
    p.xs = cstar - p.odd - 4;                               // multiple of 4 (two's complement & is the mathematical mod)
    p.ys = g.iyc - 2 * (g.R + 2) + 1;                       // row of plane (phase 0, k = 0) in the frame's buffer (guard row + 1)
    p.krows = (g.side + 4 + 3) & ~3;                        // 28, 32 or 36
    p.tsteps = (g.side + p.t0 + 1) >> 1;
    p.nblk = (p.tsteps + 3 + 3) >> 2;                       // chunks t .. t+3 of the last step
    return p;
}

// both row phases of column block b of a keypoint's patch -> slots 2b, 2b+1, completion on the block's mbarrier (one thread)
__device__ __noinline__ void issue_block(uint32_t patch_sa, uint32_t bars, const void* maps, int map, const Patch& p, int b, int fslot) {
    const uint32_t bar = bars + 8 * b;
    mbar_expect_tx(bar, 2u * (uint32_t)p.krows * 64u);
    const char* m = reinterpret_cast<const char*>(maps) + 128 * map;
    tma_load_3d(patch_sa + (2 * b) * kFSlot, m, bar, p.xs + 16 * b, p.ys, fslot);
    tma_load_3d(patch_sa + (2 * b + 1) * kFSlot, m, bar, p.xs + 16 * b, p.ys + 1, fslot);
}

template <int E, int ODD>
struct Sweep {
    // dense offsets (relative to the first column of chunk t) of the four integral columns of sample A of chunk step t
    static constexpr int kBase = 4 + ODD;
    static constexpr int kXm = kBase - 4 + E, kX0 = kBase, kX1 = kBase + 1, kXp = kBase + 5 - E;
};

}  // namespace

// ------------------------------------------------------------------------------------------ classification
// cls_cnt[slot*4 + {0,1}] = number of keypoints of the frame on the TMA path / on the gather path, cls_idx[slot][2][max_pts]
// their indices; [2], [3] are the work counters of the two kernels (all four zero on entry).
__global__ void classify_kernel(const __grid_constant__ PipeP P, const sb_point* __restrict__ points, long long pts_stride,
                                const int* __restrict__ counts, int fixed_count, int* __restrict__ cls_idx,
                                int* __restrict__ cls_cnt, int slot0, int dbg_mask) {
    pdl_wait();
    const int f = blockIdx.y, lane = threadIdx.x & 31;
    const int n = fixed_count >= 0 ? fixed_count : min(counts[f], P.max_pts);
    const sb_point* pts = points + (size_t)f * pts_stride;
    int* idx = cls_idx + (size_t)(slot0 + f) * 2 * P.max_pts;
    int* cnt = cls_cnt + (slot0 + f) * 4;
    for (int p0 = blockIdx.x * blockDim.x; p0 < n; p0 += gridDim.x * blockDim.x) {
        const int pi = p0 + threadIdx.x;
        bool fast = false;
        if (pi < n) {
            const KpGeom g = kp_geom(pts[pi].x, pts[pi].y, pts[pi].scale, P.desc_wsz, P.mag_factor, P.doubled);
            fast = fast_geom(g);
            if (dbg_mask >= 0) {  // measurement only (SB_CLS_MASK): which geometry classes leave the gather kernel's list
                const int c = g.step == 1 ? 0 : g.step == 2 ? (g.side <= kFMaxSide ? 1 : 2) : g.step == 3 ? 3 : g.step == 4 ? 4 : g.step <= 8 ? 5 : 6;
                fast = (dbg_mask >> c) & 1;
            }
        }
        const unsigned mf = __ballot_sync(0xffffffffu, pi < n && fast), ms = __ballot_sync(0xffffffffu, pi < n && !fast);
        int bf = 0, bs = 0;
        if (lane == 0) { if (mf) bf = atomicAdd(cnt + 0, __popc(mf)); if (ms) bs = atomicAdd(cnt + 1, __popc(ms)); }
        bf = __shfl_sync(0xffffffffu, bf, 0); bs = __shfl_sync(0xffffffffu, bs, 0);
        const unsigned below = (1u << lane) - 1u;
        if (pi < n) {
            if (fast) idx[bf + __popc(mf & below)] = pi;
            else idx[P.max_pts + bs + __popc(ms & below)] = pi;
        }
    }
}

// ------------------------------------------------------------------------------------------ the sweep
namespace {

struct LaneRow {            // the lattice row a lane owns
    uint32_t aM, aZ, aU, aQ;  // shared-memory address of the row in planes m = r-S, z = r, u = r+1, q = r+S+1 (block 0, chunk
                              // 0), with the row's swizzle phase already XORed into bits 4-5
    float rpos, r255;         // r255 = 0 for a row outside the window / the image: the lane computes, and adds nothing
};

// One keypoint's sweep. Branch-free inside a group of four chunk steps (eight samples): a sample's split over the cell
// columns is four weights from the column table (two of them zero), accumulated into the lane's 4 x 4 sums -- with the
// two-cell lo/hi registers of the gather kernel every sample ended in a (uniform) branch, the samples of a group could not
// overlap and the flush code made the loop body twice the size of the instruction cache's first level (ncu of the first
// version: "no instructions" was the largest stall).
template <int E, int ODD>
__device__ __forceinline__ void sweep_rows(const LaneRow& L, const Patch& pc, const float4* __restrict__ ctabW,
                                           const float* __restrict__ ctabP, uint32_t lut_sa, float* __restrict__ T, int lane,
                                           int jv_lo, int jv_hi, uint32_t bars, uint32_t& phase_bits, bool have_next,
                                           const Patch& pn, int map_next, const void* maps, uint32_t patch_sa, int fslot_next) {
    using SW = Sweep<E, ODD>;
    int4 w[4][4];  // [plane m, z, u, q][slot = chunk & 3]
    const uint32_t rowa[4] = {L.aM, L.aZ, L.aU, L.aQ};
    auto ldchunk = [&](int plane, int cidx) -> int4 {
        // chunk cidx = 4 b + cm: block b is 2 slots (two phases) further, the chunk inside the 64-byte row is swizzled
        return lds128((rowa[plane] ^ (uint32_t)((cidx & 3) << 4)) + (uint32_t)(cidx >> 2) * (2 * kFSlot));
    };
    // the next keypoint's block b replaces this keypoint's: called when every lane has consumed block b
    auto release_block = [&](int b) {
        __syncwarp();
        if (have_next && b < pn.nblk && lane == 0) issue_block(patch_sa, bars, maps, map_next, pn, b, fslot_next);
    };
    float acc[4][4];  // [cell column][sum dx, sum |dx|, sum dy, sum |dy|]
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int v = 0; v < 4; v++) acc[k][v] = 0.f;
    // Steps [tbeg, tend) in groups of four (the window's slots are static inside a group): tbeg is the first step with a
    // valid column rounded down, the last group may run up to three steps past the last valid column -- those read table
    // entries with zero weights and chunks of the buffer that nothing waits for (stale, finite integers times zero).
    const int tbeg = (jv_lo >> 1) & ~3, tend = (jv_hi + 1) >> 1;
    int waited = 0, released = 0;  // blocks [0, waited) have landed, [0, released) have been handed to the next keypoint
    auto wait_upto = [&](int b) {
#pragma unroll 1
        while (waited <= b && waited < pc.nblk) {
            mbar_wait(bars + 8 * waited, (phase_bits >> waited) & 1u);
            phase_bits ^= 1u << waited;
            waited++;
        }
    };
    wait_upto(tbeg >> 2);
#pragma unroll
    for (int s = 0; s < 4; s++)
#pragma unroll
        for (int p = 0; p < 4; p++) w[p][s] = ldchunk(p, tbeg + s);
#pragma unroll 1
    for (int tg = tbeg; tg < tend; tg += 4) {
        // blocks left of the window are free: their last chunk was consumed in the previous group
#pragma unroll 1
        while (released < (tg >> 2)) release_block(released++);
        wait_upto((tg >> 2) + 1);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int t = tg + u;
            auto val = [&](int plane, int d) -> int { return elem(w[plane][(u + (d >> 2)) & 3], d & 3); };
#pragma unroll
            for (int s2 = 0; s2 < 2; s2++) {
                const int jv = 2 * t + s2;
                const int o = 2 * s2;
                const float4 cw = ctabW[jv];  // weights of cell columns 0..3 (all 0 outside [jv_lo, jv_hi))
                const float cp2 = ctabP[jv];  // cpos^2
                // rows m (0), z (1), u (2), q (3); columns A = c-S, B = c, C = c+1, D = c+S+1
                const int mA = val(0, SW::kXm + o), mB = val(0, SW::kX0 + o), mC = val(0, SW::kX1 + o), mD = val(0, SW::kXp + o);
                const int qA = val(3, SW::kXm + o), qB = val(3, SW::kX0 + o), qC = val(3, SW::kX1 + o), qD = val(3, SW::kXp + o);
                const int zA = val(1, SW::kXm + o), zD = val(1, SW::kXp + o), uA = val(2, SW::kXm + o), uD = val(2, SW::kXp + o);
                const int wx = (qD + mB - mD - qB) - (qC + mA - mC - qA);
                const int wy = (zD - zA) + (uD - uA) - (mD - mA) - (qD - qA);
                float weight;
                asm("ld.shared.f32 %0, [%1];" : "=f"(weight) : "r"(lut_sa + 4u * (unsigned)__float2int_rz(__fmaf_rn(L.rpos, L.rpos, cp2))));
                const float a = __fmul_rn(__fmul_rn(weight, __int2float_rn(wx)), L.r255);
                const float b = __fmul_rn(__fmul_rn(weight, __int2float_rn(wy)), L.r255);
                const float cwv[4] = {cw.x, cw.y, cw.z, cw.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    acc[k][0] = __fmaf_rn(a, cwv[k], acc[k][0]); acc[k][1] = __fmaf_rn(fabsf(a), cwv[k], acc[k][1]);
                    acc[k][2] = __fmaf_rn(b, cwv[k], acc[k][2]); acc[k][3] = __fmaf_rn(fabsf(b), cwv[k], acc[k][3]);
                }
            }
            // slot u (chunk t) is dead: it takes chunk t+4, first read one step later (always inside the 5-block buffer)
#pragma unroll
            for (int p = 0; p < 4; p++) w[p][u] = ldchunk(p, t + 4);
        }
    }
    wait_upto(pc.nblk - 1);
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int v = 0; v < 4; v++) T[(k * 4 + v) * kTS + lane] = acc[k][v];
#pragma unroll 1
    while (released < pc.nblk || (have_next && released < pn.nblk)) release_block(released++);
}

}  // namespace

// grid (warps per frame, nframes), ONE warp per CTA; dynamic shared memory kFSmem + 512 (alignment)
__global__ void __launch_bounds__(32)
describe_upright_tma_kernel(const __grid_constant__ PipeP P, const void* __restrict__ maps, const sb_point* __restrict__ points,
                            long long pts_stride, const int* __restrict__ cls_idx, const int* __restrict__ cls_cnt, int slot0,
                            float* __restrict__ desc, long long desc_stride) {
    pdl_wait();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int f = blockIdx.y, lane = threadIdx.x;
    const uint32_t raw_sa = smem_u32(smem_raw);
    const uint32_t patch_sa = (raw_sa + 511u) & ~511u;
    unsigned char* base = smem_raw + (patch_sa - raw_sa);
    float* T = reinterpret_cast<float*>(base + kFPatch);
    float* Wr = T + 16 * kTS;
    float4* ctabW = reinterpret_cast<float4*>(Wr + 4 * kTS);
    float* ctabP = reinterpret_cast<float*>(ctabW + kFColTab);
    float* s_lut2 = ctabP + kFColTab;
    const uint32_t bars = smem_u32(s_lut2 + 40);
    for (int t = lane; t < 40; t += 32) s_lut2[t] = P.lut2[t];
    // the table entries past the 32 virtual columns a keypoint can have stay zero: the last group of a sweep may touch them
    for (int t = lane; t < kFColTab; t += 32) { ctabW[t] = make_float4(0.f, 0.f, 0.f, 0.f); ctabP[t] = 0.f; }
    if (lane == 0) {
        for (int b = 0; b < kFNB; b++) mbar_init(bars + 8 * b, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const uint32_t lut_sa = smem_u32(s_lut2);
    const int* list = cls_idx + (size_t)(slot0 + f) * 2 * P.max_pts;
    const int n = cls_cnt[(slot0 + f) * 4];
    const sb_point* pts = points + (size_t)f * pts_stride;
    float* dout = desc + (size_t)f * desc_stride;
    const int W = 4;
    const float fW = 4.f, wofs = 1.5f;
    uint32_t phase_bits = 0;

    // keypoints cur, cur + gridDim.x, ... of the frame's list (they cost within +-30 % of each other on this path, and a
    // dynamic counter would put an atomic round trip in front of every prefetch)
    int cur = blockIdx.x;
    if (cur >= n) return;
    int pi = list[cur];
    KpGeom kg = kp_geom(pts[pi].x, pts[pi].y, pts[pi].scale, W, P.mag_factor, P.doubled);
    Patch pc = patch_of(kg);
    if (lane == 0)  // the first keypoint's patch: all blocks at once
        for (int b = 0; b < pc.nblk; b++) issue_block(patch_sa, bars, maps, (pc.krows >> 2) - 7, pc, b, slot0 + f);
    while (true) {
        // the keypoint after this one: its patch follows this one's through the buffer
        const int nxt = cur + gridDim.x;
        const bool have_next = nxt < n;
        int pin = 0;
        KpGeom kn = kg;
        Patch pn = pc;
        if (have_next) {
            pin = list[nxt];
            kn = kp_geom(pts[pin].x, pts[pin].y, pts[pin].scale, W, P.mag_factor, P.doubled);
            pn = patch_of(kn);
        }
        const int step = kg.step, ixc = kg.ixc, iyc = kg.iyc, S = kg.S, R = kg.R, side = kg.side;
        const float fx = kg.fx, fy = kg.fy, spacing = kg.spacing;
        // ---- column table: lane = virtual column jv (side + 1 <= 32)
        int jv_lo = 32, jv_hi = 0;
        {
            const int jj = lane - pc.t0;
            const int j = jj - R;
            const float cpos = __fdiv_rn(__fsub_rn(__int2float_rn(step * j), fx), spacing);
            const float cx = __fadd_rn(cpos, wofs);
            const int c = ixc + j * step;
            const bool ok = jj >= 0 && jj < side && cx > -1.f && cx < fW && c >= 1 + S && c < P.iw - 1 - S;
            const int ci = __float2int_rz(cx >= 0.f ? cx : __fsub_rn(cx, 1.f));
            const float cfrac = __fsub_rn(cx, __int2float_rn(ci)), cfrac1 = __fsub_rn(1.f, cfrac);
            float4 cw;
            cw.x = ok ? ((ci == 0 ? cfrac1 : 0.f) + (ci == -1 ? cfrac : 0.f)) : 0.f;
            cw.y = ok ? ((ci == 1 ? cfrac1 : 0.f) + (ci == 0 ? cfrac : 0.f)) : 0.f;
            cw.z = ok ? ((ci == 2 ? cfrac1 : 0.f) + (ci == 1 ? cfrac : 0.f)) : 0.f;
            cw.w = ok ? ((ci == 3 ? cfrac1 : 0.f) + (ci == 2 ? cfrac : 0.f)) : 0.f;
            ctabW[lane] = cw;
            ctabP[lane] = ok ? __fmul_rn(cpos, cpos) : 0.f;
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (m) { jv_lo = __ffs(m) - 1; jv_hi = 32 - __clz(m); }
        }
        // ---- the lane's row
        LaneRow L;
        float rw0 = 0.f, rw1 = 0.f;
        int ri = -8;
        {
            const int ii = min(lane, side - 1);
            const int i = ii - R;
            const float rpos = __fdiv_rn(__fsub_rn(__int2float_rn(step * i), fy), spacing);
            const float rx = __fadd_rn(rpos, wofs);
            const int r = iyc + i * step;
            const bool ok = lane < side && rx > -1.f && rx < fW && r >= 1 + S && r < P.ih - 1 - S;
            L.rpos = ok ? rpos : 0.f;
            L.r255 = ok ? kR255 : 0.f;
            if (ok) {
                ri = __float2int_rz(rx >= 0.f ? rx : __fsub_rn(rx, 1.f));
                rw1 = __fsub_rn(rx, __int2float_rn(ri));
                rw0 = __fsub_rn(1.f, rw1);
            }
            const int k = ii + 2;
            const int e = kg.e;
            // plane (phase a, row k): slot offset a * kFSlot, row k * 64, swizzle phase ((k >> 1) & 3) ^ (a << 1) in bits 4-5
            auto rowaddr = [&](int a, int kk) -> uint32_t {
                return patch_sa + (uint32_t)(a * kFSlot + kk * 64) + (uint32_t)(((((kk >> 1) & 3) ^ (a << 1)) & 3) << 4);
            };
            L.aM = rowaddr(e, k - 2);
            L.aZ = rowaddr(0, k);
            L.aU = rowaddr(1, k);
            L.aQ = rowaddr(1 - e, k + 2);
        }
#pragma unroll
        for (int kr = 0; kr < 4; kr++) Wr[kr * kTS + lane] = (ri == kr ? rw0 : 0.f) + (ri + 1 == kr ? rw1 : 0.f);
        __syncwarp();
        float acc[2] = {0.f, 0.f};
        if (jv_hi > jv_lo) {
            const int mapn = (pn.krows >> 2) - 7;
            if (kg.e == 0) {
                if (pc.odd) sweep_rows<0, 1>(L, pc, ctabW, ctabP, lut_sa, T, lane, jv_lo, jv_hi, bars, phase_bits, have_next, pn, mapn, maps, patch_sa, slot0 + f);
                else sweep_rows<0, 0>(L, pc, ctabW, ctabP, lut_sa, T, lane, jv_lo, jv_hi, bars, phase_bits, have_next, pn, mapn, maps, patch_sa, slot0 + f);
            } else {
                if (pc.odd) sweep_rows<1, 1>(L, pc, ctabW, ctabP, lut_sa, T, lane, jv_lo, jv_hi, bars, phase_bits, have_next, pn, mapn, maps, patch_sa, slot0 + f);
                else sweep_rows<1, 0>(L, pc, ctabW, ctabP, lut_sa, T, lane, jv_lo, jv_hi, bars, phase_bits, have_next, pn, mapn, maps, patch_sa, slot0 + f);
            }
            __syncwarp();
            // element e = (kr*4 + kc)*4 + o gathers sum_l Wr[kr][l] * T[kc*4 + o][l] over the 32 lanes' rows
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int e = lane + 32 * u;
                const int o = e & 3, t = e >> 2, kr = t >> 2, kc = t & 3;
                const float4* tr = reinterpret_cast<const float4*>(T + (kc * 4 + o) * kTS);
                const float4* wr = reinterpret_cast<const float4*>(Wr + kr * kTS);
                float a = 0.f;
#pragma unroll
                for (int l4 = 0; l4 < 8; l4++) {
                    const float4 tv = tr[l4], wv = wr[l4];
                    a = __fmaf_rn(wv.x, tv.x, a); a = __fmaf_rn(wv.y, tv.y, a);
                    a = __fmaf_rn(wv.z, tv.z, a); a = __fmaf_rn(wv.w, tv.w, a);
                }
                acc[u] = a;
            }
        } else {
            // no valid column: nothing was read, but the patch still has to make room for the next keypoint's
            __syncwarp();
            for (int b = 0; b < pc.nblk; b++) { mbar_wait(bars + 8 * b, (phase_bits >> b) & 1u); phase_bits ^= 1u << b; }
            if (have_next && lane == 0)
                for (int b = 0; b < pn.nblk; b++) issue_block(patch_sa, bars, maps, (pn.krows >> 2) - 7, pn, b, slot0 + f);
        }
        // split by sign (elements o = 0,1 hold S, A of dx and become (S-A)/2, (S+A)/2; o = 2,3 likewise for dy), normalise, store
        float v[2];
        float sq = 0.f;
#pragma unroll
        for (int u = 0; u < 2; u++) {
            float a = acc[u];
            const float pr = __shfl_xor_sync(0xffffffffu, a, 1);
            a = (lane & 1) ? 0.5f * (pr + a) : 0.5f * (a - pr);
            v[u] = a;
            sq = __fmaf_rn(a, a, sq);
        }
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o2);
        const float inv = __fdiv_rn(1.f, __fsqrt_rn(sq));
        float* d = dout + (size_t)pi * 64;
        d[lane] = __fmul_rn(v[0], inv);
        d[lane + 32] = __fmul_rn(v[1], inv);
        __syncwarp();
        if (!have_next) break;
        cur = nxt; pi = pin; kg = kn; pc = pn;
    }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// three tensor maps over the context's integral buffer [batch][ih + 2][ip] int32: boxes of 16 columns x K rows taken every
// SECOND row (elementStrides {1, 2, 1}), K = 28, 32, 36, 64-byte swizzle; copied to a 128-byte-strided device array
cudaError_t build_describe_maps(const PipeP& P, const int* d_integral, int batch, void** d_maps) {
    *d_maps = nullptr;
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    if (e != cudaSuccess) return e;
    if (q != cudaDriverEntryPointSuccess || !fp) return cudaErrorNotSupported;
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fp);
    alignas(64) CUtensorMap maps[3];
    static_assert(sizeof(CUtensorMap) == 128, "tensor map stride");
    for (int m = 0; m < 3; m++) {
        const int K = 28 + 4 * m;
        const cuuint64_t dims[3] = {(cuuint64_t)P.ip, (cuuint64_t)(P.ih + 2), (cuuint64_t)batch};
        const cuuint64_t strides[2] = {(cuuint64_t)P.ip * 4, (cuuint64_t)P.istride * 4};
        const cuuint32_t box[3] = {16, (cuuint32_t)(2 * K), 1};
        const cuuint32_t estr[3] = {1, 2, 1};
        if (enc(&maps[m], CU_TENSOR_MAP_DATA_TYPE_INT32, 3, const_cast<int*>(d_integral), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return cudaErrorInvalidValue;
    }
    if ((e = cudaMalloc(d_maps, sizeof(maps))) != cudaSuccess) return e;
    return cudaMemcpy(*d_maps, maps, sizeof(maps), cudaMemcpyHostToDevice);
}

bool describe_tma_applies(const PipeP& P) { return P.upright && P.orient_size == 4 && P.desc_wsz == 4; }

cudaError_t launch_describe_tma(const PipeP& P, int nframes, const DescAux& aux, sb_point* d_points, long long pts_stride,
                                const int* d_counts, int fixed_count, float* d_desc, long long desc_stride, int sm_count,
                                cudaStream_t st) {
    const int maxn = fixed_count >= 0 ? fixed_count : P.max_pts;
    static const int dbg_mask = getenv("SB_CLS_MASK") ? atoi(getenv("SB_CLS_MASK")) : -1;  // measurement only: skips the TMA kernel
    cudaError_t e = launch_dep(classify_kernel, dim3(max(1, min(16, (maxn + 255) / 256)), nframes), dim3(256), 0, st, P, d_points, pts_stride,
                               d_counts, fixed_count, aux.cls_idx, aux.cls_cnt, aux.slot0, dbg_mask);
    if (e != cudaSuccess || dbg_mask >= 0) return e;
    const int smem = kFSmem + 512;
    e = cudaFuncSetAttribute(describe_upright_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    const int resident = sm_count * (int)((227 * 1024) / (smem + 1024));
    int warps = max(1, resident / nframes);
    if (warps > maxn) warps = maxn;
    return launch_dep(describe_upright_tma_kernel, dim3(warps, nframes), dim3(32), smem, st, P, aux.maps, d_points, pts_stride,
                      (const int*)aux.cls_idx, (const int*)aux.cls_cnt, aux.slot0, d_desc, desc_stride);
}

}  // namespace sb
