// nms.cu -- 3x3x3 scale-space non-maximum suppression, sub-pixel quadratic refinement and keypoint
// append, for all octaves of a frame batch in ONE launch.
//
// Replaces cuFindMaximumWithInterp / findMaximumWithInterp + fitQuadrat + solveLinearSystem +
// makePoint (surfd.cu:676-832, 835-887, 942-988, 1001-1022, 3058-3079), which the reference launches
// once per octave after two synchronous cudaMemcpyToSymbol. Same decomposition into 2x2x2 cells
// (equivalent to 3x3x3 NMS with the reference's tie rules: strict > inside the cell, reject only
// if best < neighbour outside it). Differences by design:
//   * keypoints are appended with ONE atomicAdd per warp (ballot + popc), bounded by max_pts --
//     the reference does one atomicInc per keypoint and can write past the buffer;
//   * SurfPoint.o is set to the octave index (the reference never writes it);
//   * no global descriptor-radius atomicMax (the describe kernel sizes its own sampling lattice).
#include "common.cuh"

namespace sb {

// 3x3 Gaussian elimination with partial pivoting (surfd.cu:835-887); `a -= m*b` as one FFMA and
// IEEE division, as in the reference's SASS.
__device__ __forceinline__ void solve3(float* sol, float (&sq)[3][3]) {
    int pivot = 0;
#pragma unroll
    for (int col = 0; col < 2; col++) {
        float maxc = -1.f;
#pragma unroll
        for (int row = col; row < 3; row++) {
            const float coef = fabsf(sq[row][col]);
            if (coef > maxc) { maxc = coef; pivot = row; }
        }
        if (pivot != col) {
#pragma unroll
            for (int i = 0; i < 3; i++) {
                // pivot is a runtime index: select without dynamic register indexing
                float pv = (pivot == 1) ? sq[1][i] : sq[2][i];
                const float cv = sq[col][i];
                if (pivot == 1) sq[1][i] = cv; else sq[2][i] = cv;
                sq[col][i] = pv;
            }
            float ps = (pivot == 1) ? sol[1] : sol[2];
            const float cs = sol[col];
            if (pivot == 1) sol[1] = cs; else sol[2] = cs;
            sol[col] = ps;
        }
#pragma unroll
        for (int row = col + 1; row < 3; row++) {
            const float mult = __fdiv_rn(sq[row][col], sq[col][col]);
#pragma unroll
            for (int c = col; c < 3; c++) sq[row][c] = __fmaf_rn(-mult, sq[col][c], sq[row][c]);
            sol[row] = __fmaf_rn(-mult, sol[col], sol[row]);
        }
    }
#pragma unroll
    for (int row = 2; row >= 0; row--) {
        float val = sol[row];
#pragma unroll
        for (int col = 2; col > row; col--) val = __fmaf_rn(-sol[col], sq[row][col], val);
        sol[row] = __fdiv_rn(val, sq[row][row]);
    }
}

// Quadratic fit around (s,r,c): gradient and Hessian by central differences (surfd.cu:942-988).
// Returns the interpolated strength; off[] = -H^-1 g.
__device__ __forceinline__ float fit_quadratic(const float* __restrict__ src, int sp, int osz, int s, int r, int c,
                                               float (&off)[3]) {
    const float* cur = src + (size_t)s * osz + (size_t)r * sp + c;
    const float* prv = cur - osz;
    const float* nxt = cur + osz;
    const float v = __ldg(cur);
    const float cn = __ldg(cur + sp), cp = __ldg(cur - sp), ce = __ldg(cur + 1), cw = __ldg(cur - 1);
    const float pv = __ldg(prv), nv = __ldg(nxt);
    float g[3], H[3][3];
    g[0] = __fmul_rn(__fsub_rn(nv, pv), 0.5f);
    g[1] = __fmul_rn(__fsub_rn(cn, cp), 0.5f);
    g[2] = __fmul_rn(__fsub_rn(ce, cw), 0.5f);
    const float temp = __fadd_rn(v, v);
    H[0][0] = __fsub_rn(__fadd_rn(pv, nv), temp);
    H[1][1] = __fsub_rn(__fadd_rn(cn, cp), temp);
    H[2][2] = __fsub_rn(__fadd_rn(ce, cw), temp);
    H[0][1] = __fmul_rn(__fsub_rn(__fsub_rn(__ldg(nxt + sp), __ldg(nxt - sp)), __fsub_rn(__ldg(prv + sp), __ldg(prv - sp))), 0.25f);
    H[0][2] = __fmul_rn(__fsub_rn(__fsub_rn(__ldg(nxt + 1), __ldg(nxt - 1)), __fsub_rn(__ldg(prv + 1), __ldg(prv - 1))), 0.25f);
    H[1][2] = __fmul_rn(__fsub_rn(__fsub_rn(__ldg(cur + sp + 1), __ldg(cur + sp - 1)),
                                  __fsub_rn(__ldg(cur - sp + 1), __ldg(cur - sp - 1))), 0.25f);
    H[1][0] = H[0][1];
    H[2][0] = H[0][2];
    H[2][1] = H[1][2];
    off[0] = -g[0];
    off[1] = -g[1];
    off[2] = -g[2];
    solve3(off, H);
    // reference SASS: t = o1*g1; t = fma(o0,g0,t); t = fma(o2,g2,t); strength = fma(t, 0.5, v)
    float t = __fmul_rn(off[1], g[1]);
    t = __fmaf_rn(off[0], g[0], t);
    t = __fmaf_rn(off[2], g[2], t);
    return __fmaf_rn(t, 0.5f, v);
}

// sign of the Laplacian at the keypoint's own lobe (getTrace, surfd.cu:369-377)
__device__ __forceinline__ int laplace_sign(const int* __restrict__ I, int ip, int cx, int cy, int l) {
    const int x2 = l / 2, x3 = x2 + x2;
    const int lxx = box_sum(I, ip, cx - l - x2, cx + l + x2, cy - x3, cy + x3) - 3 * box_sum(I, ip, cx - x2, cx + x2, cy - x3, cy + x3);
    const int lyy = box_sum(I, ip, cx - x3, cx + x3, cy - l - x2, cy + l + x2) - 3 * box_sum(I, ip, cx - x3, cx + x3, cy - x2, cy + x2);
    return (lxx + lyy > 0) ? 1 : -1;
}

// Two kernels. The scan touches every response value once and must run at memory speed, the refinement runs for
// ~1.5 % of the cells and needs 60 registers; fused, the refinement's registers cut the scan's occupancy to a third
// (ncu, fused version: 58 registers, 34 % warps active, 25 % of DRAM peak).
//   nms_scan_kernel   thread per 2x2x2 cell: cell maximum, 0.8*thresh test, top-layer rule, 19-neighbour test;
//                     survivors go to a per-frame candidate queue as one packed word (octave, layer, row, column)
//   nms_refine_kernel thread per candidate: <= 5 quadratic fits, rejection tests, makePoint, append
constexpr int kCandShiftS = 26, kCandShiftO = 29;  // packed candidate: c [0,13) | r [13,26) | s [26,29) | o [29,32)

// grid (nms_tiles, nframes), block 32x8: one thread per 2x2(xy) x 2(scale) cell.
// (A warp-cooperative neighbour test -- lane n loads neighbour n of one candidate, one ballot decides -- and four
// cells per thread were tried: 2 % faster in a 64-frame batch, 40 % slower for a single frame, because a warp's
// candidates then resolve one L2 round trip after the other. Not kept.)
__global__ void __launch_bounds__(256)
nms_scan_kernel(const __grid_constant__ PipeP P, const float* __restrict__ Rbase, unsigned* __restrict__ cand,
                int* __restrict__ cand_count, int cand_cap) {
    pdl_wait();
    const int f = blockIdx.y;
    const int tile = blockIdx.x;
    int o = 0;
    while (o + 1 < P.noctaves && tile >= P.oct[o + 1].nms_tile0) o++;
    const OctaveP& q = P.oct[o];
    int lt = tile - q.nms_tile0;
    const int per_z = q.nms_tx * q.nms_ty;
    int z = 0;
    while (lt >= per_z) { lt -= per_z; z++; }  // z < nmb <= 3
    const int ty = div_small(lt, q.inv_nms_tx), tx = lt - ty * q.nms_tx;
    const int lane = threadIdx.x;
    const int xc = tx * 32 + lane, yc = ty * 8 + threadIdx.y;

    const int ms = P.max_scale;
    const int k = 2 * z + 1;
    const int mb = q.mb[z];
    const int i = mb + 2 * yc, j = mb + 2 * xc;
    const int sw = q.sw, sh = q.sh, sp = q.sp, osz = q.osz;
    const float* src = Rbase + (size_t)f * P.rstride + q.resp_off;
    asm volatile("" : "+l"(src));  // one materialised base pointer: addresses are IMAD.WIDE(index, 4, src)

    bool cand_ok = false;
    unsigned packed = 0;
    if (k < ms - 1 && i < sh - mb && j < sw - mb) {
        // UNSIGNED 32-bit element indices from one materialised base pointer (an octave's layers are < 2^31 elements and
        // every index is >= 0): an address is IADD + IMAD.WIDE.U32; with size_t products -- or signed ints, which the
        // compiler widens before adding -- the address arithmetic was half of the kernel's instructions
        const unsigned usp = sp, uosz = osz;
        const unsigned u0 = (unsigned)k * uosz + (unsigned)i * usp + (unsigned)j, u1 = u0 + uosz;
        // cell maximum in the reference's scan order, strict > (surfd.cu:699-736)
        const float *pa = src + u0, *pb = src + (u0 + usp), *pc = src + u1, *pd = src + (u1 + usp);  // +-1 column: immediate offsets
        const float v0 = __ldg(pa), v1 = __ldg(pa + 1), v2 = __ldg(pb), v3 = __ldg(pb + 1);
        const float v4 = __ldg(pc), v5 = __ldg(pc + 1), v6 = __ldg(pd), v7 = __ldg(pd + 1);
        float best = v0;
        int cas = 0;
        if (v1 > best) { best = v1; cas = 1; }
        if (v2 > best) { best = v2; cas = 2; }
        if (v3 > best) { best = v3; cas = 3; }
        if (v4 > best) { best = v4; cas = 4; }
        if (v5 > best) { best = v5; cas = 5; }
        if (v6 > best) { best = v6; cas = 6; }
        if (v7 > best) { best = v7; cas = 7; }
        bool cnd = !(best < __fmul_rn(P.thresh, 0.8f)) && !(k + 1 == ms - 1 && cas > 3);
        const int s = k + (cas >> 2), r = i + ((cas >> 1) & 1), c = j + (cas & 1);
        if (cnd) {
            // outward directions: the cell's other member along each axis sits at -d (two's-complement offsets)
            const unsigned dso = (cas & 4) ? uosz : 0u - uosz, drp = (cas & 2) ? usp : 0u - usp, dc = (cas & 1) ? 1u : ~0u;
            const unsigned ci = (unsigned)s * uosz + (unsigned)r * usp + (unsigned)c;
            // outer layer s+ds: all nine
            {
                const unsigned l1 = ci + dso;
                const float *q0 = src + (l1 - usp), *q1 = src + l1, *q2 = src + (l1 + usp);
                if (best < __ldg(q0 - 1)) cnd = false;
                if (best < __ldg(q0)) cnd = false;
                if (best < __ldg(q0 + 1)) cnd = false;
                if (best < __ldg(q1 - 1)) cnd = false;
                if (best < __ldg(q1)) cnd = false;
                if (best < __ldg(q1 + 1)) cnd = false;
                if (best < __ldg(q2 - 1)) cnd = false;
                if (best < __ldg(q2)) cnd = false;
                if (best < __ldg(q2 + 1)) cnd = false;
            }
            // own layer s and inner layer s-ds: the five positions outside the cell's 2x2 footprint
#pragma unroll
            for (int li = 0; li < 2; li++) {
                const unsigned m = li ? ci - dso : ci;
                const float* rowo = src + (m + drp);  // outward row: three
                if (best < __ldg(rowo - 1)) cnd = false;
                if (best < __ldg(rowo)) cnd = false;
                if (best < __ldg(rowo + 1)) cnd = false;
                const unsigned side = m + dc;   // (r, c+dc) and (r-dr, c+dc)
                if (best < __ldg(src + side)) cnd = false;
                if (best < __ldg(src + (side - drp))) cnd = false;
            }
        }
        cand_ok = cnd;
        packed = (unsigned)c | ((unsigned)r << 13) | ((unsigned)s << kCandShiftS) | ((unsigned)o << kCandShiftO);
    }
    const unsigned m = __ballot_sync(0xffffffffu, cand_ok);
    if (m) {
        const int leader = __ffs(m) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(&cand_count[f], __popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (cand_ok) {
            const int slot = base + __popc(m & ((1u << lane) - 1u));
            if (slot < cand_cap) cand[(size_t)f * cand_cap + slot] = packed;
        }
    }
}

// The same scan for the standard 5-layer octave (max_scale 5: cell layers z = 0, 1 <-> k = 1, 3), staged through shared
// memory. The per-cell kernel above makes two dependent trips to memory (the cell's 8 values, then -- for the warps that
// hold a candidate, i.e. most of them -- its 19 neighbours) and reads every layer from two z-slices; ncu showed it
// latency-bound (long scoreboard 11 per issue, 36 % of DRAM peak). Here a CTA loads the 5 layers under its 32x8 cells
// (18+dm rows x 66+dm columns, dm = offset between the two cell lattices) once, all loads in flight together, and both
// z-slices are decided from shared memory. grid (sum over octaves of nms_tx*nms_ty, nframes), block 32x8.
constexpr int kNmsR = 20, kNmsC = 72;  // staged rows / padded columns: dm <= 2

__global__ void __launch_bounds__(256, 4)
nms_scan_tile_kernel(const __grid_constant__ PipeP P, const float* __restrict__ Rbase, unsigned* __restrict__ cand,
                     int* __restrict__ cand_count, int cand_cap) {
    pdl_wait();
    __shared__ __align__(16) float blk[5][kNmsR][kNmsC];
    const int f = blockIdx.y;
    int lt = blockIdx.x, o = 0;
    while (o + 1 < P.noctaves && lt >= P.oct[o].nms_tx * P.oct[o].nms_ty) { lt -= P.oct[o].nms_tx * P.oct[o].nms_ty; o++; }
    const OctaveP& q = P.oct[o];
    const int ty = div_small(lt, q.inv_nms_tx), tx = lt - ty * q.nms_tx;
    const int mlo = min(q.mb[0], q.mb[1]), dm = max(q.mb[0], q.mb[1]) - mlo;
    // staged block: rows row0 .. row0+NR-1, columns from col0 rounded down to a multiple of 4 (128-bit loads; the response
    // rows are 128-byte aligned and zero beyond their width), 72 columns = 66 + dm + the rounding
    const int row0 = mlo + 16 * ty - 1, col0 = (mlo + 64 * tx - 1) & ~3;
    const int NR = 18 + dm;
    const int sw = q.sw, sh = q.sh;
    const unsigned usp = q.sp, uosz = q.osz;
    const float* src = Rbase + (size_t)f * P.rstride + q.resp_off;
    asm volatile("" : "+l"(src));
    const int lane = threadIdx.x, ly = threadIdx.y, tid = ly * 32 + lane;
    {
        // 5 layers x 20 rows x 18 float4 = 1800 positions, 8 per thread, all 8 x 128-bit loads in flight before the first
        // shared-memory store waits on one
        float4* flat = reinterpret_cast<float4*>(&blk[0][0][0]);
        float4 v[8];
        int si[8];
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const int e = tid + 256 * it;
            const int rr = e / (kNmsC / 4), c4 = e - rr * (kNmsC / 4);  // rr = layer * kNmsR + row
            const int layer = rr / kNmsR, r = rr - layer * kNmsR;
            const int gr = row0 + r, gc = col0 + 4 * c4;
            const bool in = layer < 5 && r < NR && gr < sh && gc < (int)usp;
            si[it] = (layer < 5 && r < NR) ? e : -1;
            // clamped address + select instead of a branch around the load
            const unsigned gi = in ? (unsigned)layer * uosz + (unsigned)gr * usp + (unsigned)gc : 0u;
            const float4 t = __ldg(reinterpret_cast<const float4*>(src + gi));
            v[it] = in ? t : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int it = 0; it < 8; it++)
            if (si[it] >= 0) flat[si[it]] = v[it];
    }
    __syncthreads();

    const int ms = P.max_scale;
    constexpr int LS = kNmsR * kNmsC;  // layer stride in the staged block
    bool okz[2] = {false, false};
    unsigned pkz[2] = {0u, 0u};
#pragma unroll
    for (int z = 0; z < 2; z++) {
        const int k = 2 * z + 1;
        const int mb = q.mb[z];
        const int i = mb + 2 * (8 * ty + ly), j = mb + 2 * (32 * tx + lane);
        if (i < sh - mb && j < sw - mb) {
            const float* c0 = &blk[k][i - row0][j - col0];
            const float* c1 = c0 + LS;
            // cell maximum in the reference's scan order, strict > (surfd.cu:699-736)
            const float v0 = c0[0], v1 = c0[1], v2 = c0[kNmsC], v3 = c0[kNmsC + 1];
            const float v4 = c1[0], v5 = c1[1], v6 = c1[kNmsC], v7 = c1[kNmsC + 1];
            float best = v0;
            int cas = 0;
            if (v1 > best) { best = v1; cas = 1; }
            if (v2 > best) { best = v2; cas = 2; }
            if (v3 > best) { best = v3; cas = 3; }
            if (v4 > best) { best = v4; cas = 4; }
            if (v5 > best) { best = v5; cas = 5; }
            if (v6 > best) { best = v6; cas = 6; }
            if (v7 > best) { best = v7; cas = 7; }
            bool cnd = !(best < __fmul_rn(P.thresh, 0.8f)) && !(k + 1 == ms - 1 && cas > 3);
            const int s = k + (cas >> 2), r = i + ((cas >> 1) & 1), c = j + (cas & 1);
            if (cnd) {
                // outward directions: the cell's other member along each axis sits at -d
                const int dso = (cas & 4) ? LS : -LS, drp = (cas & 2) ? kNmsC : -kNmsC, dc = (cas & 1) ? 1 : -1;
                const float* ctr = &blk[s][r - row0][c - col0];
                const float* q1 = ctr + dso;  // outer layer s+ds: all nine
                if (best < q1[-kNmsC - 1]) cnd = false;
                if (best < q1[-kNmsC]) cnd = false;
                if (best < q1[-kNmsC + 1]) cnd = false;
                if (best < q1[-1]) cnd = false;
                if (best < q1[0]) cnd = false;
                if (best < q1[1]) cnd = false;
                if (best < q1[kNmsC - 1]) cnd = false;
                if (best < q1[kNmsC]) cnd = false;
                if (best < q1[kNmsC + 1]) cnd = false;
                // own layer s and inner layer s-ds: the five positions outside the cell's 2x2 footprint
#pragma unroll
                for (int li = 0; li < 2; li++) {
                    const float* m = li ? ctr - dso : ctr;
                    const float* rowo = m + drp;  // outward row: three
                    if (best < rowo[-1]) cnd = false;
                    if (best < rowo[0]) cnd = false;
                    if (best < rowo[1]) cnd = false;
                    if (best < m[dc]) cnd = false;        // (r, c+dc)
                    if (best < m[dc - drp]) cnd = false;  // (r-dr, c+dc)
                }
            }
            okz[z] = cnd;
            pkz[z] = (unsigned)c | ((unsigned)r << 13) | ((unsigned)s << kCandShiftS) | ((unsigned)o << kCandShiftO);
        }
    }
    const unsigned m0 = __ballot_sync(0xffffffffu, okz[0]), m1 = __ballot_sync(0xffffffffu, okz[1]);
    if (m0 | m1) {
        const int n0 = __popc(m0);
        int base = 0;
        if (lane == 0) base = atomicAdd(&cand_count[f], n0 + __popc(m1));
        base = __shfl_sync(0xffffffffu, base, 0);
        const unsigned below = (1u << lane) - 1u;
        if (okz[0]) {
            const int slot = base + __popc(m0 & below);
            if (slot < cand_cap) cand[(size_t)f * cand_cap + slot] = pkz[0];
        }
        if (okz[1]) {
            const int slot = base + n0 + __popc(m1 & below);
            if (slot < cand_cap) cand[(size_t)f * cand_cap + slot] = pkz[1];
        }
    }
}

// grid (ctas, nframes), 128 threads, grid-stride over the frame's candidates.
__global__ void __launch_bounds__(128)
nms_refine_kernel(const __grid_constant__ PipeP P, const int* __restrict__ Ibase, const float* __restrict__ Rbase,
                  const unsigned* __restrict__ cand, int* cand_count, int cand_cap,
                  sb_point* __restrict__ points, int* counts, unsigned* __restrict__ done, int* work, int* work2, int* cls_cnt) {
    pdl_wait();
    const int f = blockIdx.y, lane = threadIdx.x & 31;
    const int ncand = min(cand_count[f], cand_cap);
    const int ms = P.max_scale;
    // whole warps iterate together (the append below is warp-aggregated)
    for (int base_i = blockIdx.x * blockDim.x + (threadIdx.x & ~31); base_i < ncand; base_i += gridDim.x * blockDim.x) {
        const int ci = base_i + lane;
        bool keep = false;
        float kx = 0.f, ky = 0.f, kscale = 0.f, kstrength = 0.f;
        int klap = 1, o = 0;
        if (ci < ncand) {
            const unsigned pk = cand[(size_t)f * cand_cap + ci];
            o = (int)(pk >> kCandShiftO);
            const int s = (int)((pk >> kCandShiftS) & 7u);
            int r = (int)((pk >> 13) & 0x1fffu), c = (int)(pk & 0x1fffu);
            const OctaveP& q = P.oct[o];
            const int sw = q.sw, sh = q.sh, sp = q.sp, osz = q.osz;
            const float* src = Rbase + (size_t)f * P.rstride + q.resp_off;
            float off[3] = {0.f, 0.f, 0.f};
            float strength = 0.f;
            int newr = r, newc = c;
            const int bs = q.borders[s];
            for (int mv = 0; mv < 5; mv++) {
                r = newr; c = newc;
                strength = fit_quadratic(src, sp, osz, s, r, c, off);
                if (off[1] > 0.6f && r < sh - bs) newr++;
                if (off[1] < -0.6f && r > bs) newr--;
                if (off[2] > 0.6f && c < sw - bs) newc++;
                if (off[2] < -0.6f && c > bs) newc--;
                if (newr == r && newc == c) break;
            }
            const bool bad = isnan(off[0]) || isnan(off[1]) || isnan(off[2]) || fabsf(off[0]) > 1.5f ||
                             fabsf(off[1]) > 1.5f || fabsf(off[2]) > 1.5f || strength < P.thresh;
            if (!bad) {
                const int octave = q.octave;
                // ns = ((s+off0)*2*octave + init_lobe + (octave-1)*max_scale) / 3   (surfd.cu:822)
                float t = __fadd_rn(__int2float_rn(s), off[0]);
                t = __fadd_rn(t, t);
                const float ns = __fdiv_rn(__fmaf_rn(t, __int2float_rn(octave), __int2float_rn(P.init_lobe + (octave - 1) * ms)), 3.f);
                const float ny = __fmul_rn(__int2float_rn(octave), __fadd_rn(__int2float_rn(r), off[1]));
                const float nx = __fmul_rn(__int2float_rn(octave), __fadd_rn(__int2float_rn(c), off[2]));
                // makePoint (surfd.cu:1001-1022)
                const float fs = __int2float_rn(P.sampling);
                const float td = __fmul_rn(fs, P.divisor);
                kx = __fmul_rn(nx, td);
                ky = __fmul_rn(ny, td);
                kscale = __fmul_rn(__fmul_rn(1.2f, ns), P.divisor);
                kstrength = strength;
                const int lobe = __float2int_rz(__fmaf_rn(3.f, ns, 0.5f));
                const int px = __float2int_rz(__fmaf_rn(nx, fs, 0.5f));
                const int py = __float2int_rz(__fmaf_rn(ny, fs, 0.5f));
                klap = laplace_sign(Ibase + (size_t)f * P.istride + P.ip, P.ip, px, py, lobe);
                keep = true;
            }
        }
        // warp-aggregated append: one atomic per warp, hard bound on the slot
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (m) {
            const int leader = __ffs(m) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(&counts[f], __popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (keep) {
                const int slot = base + __popc(m & ((1u << lane) - 1u));
                if (slot < P.max_pts) {
                    // 48-byte SurfPoint as three 128-bit stores
                    float4* dst = reinterpret_cast<float4*>(points + (size_t)f * P.max_pts + slot);
                    dst[0] = make_float4(kx, ky, kscale, __int_as_float(o));
                    dst[1] = make_float4(kstrength, __int_as_float(klap), 0.f /*ori*/, 0.f /*score*/);
                    dst[2] = make_float4(__int_as_float(-1) /*match*/, 0.f, 0.f, 0.f);
                }
            }
        }
    }
    // The block that finishes a frame's refinement last closes the frame: the keypoint count is clamped to the capacity and the
    // candidate counter and the work counters of the descriptor kernels are re-armed for the next call (they are zero after
    // sb_create, and every pass leaves them zero again: no memset and no extra launch in the per-frame sequence). atomicInc
    // wraps at gridDim.x - 1, so `done` re-arms itself.
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicInc(done + f, gridDim.x - 1) == gridDim.x - 1) {
            __threadfence();
            const int c = *reinterpret_cast<volatile int*>(counts + f);
            counts[f] = min(c, P.max_pts);
            cand_count[f] = 0; work[f] = 0; work2[f] = 0;
            if (cls_cnt) *reinterpret_cast<int4*>(cls_cnt + 4 * f) = make_int4(0, 0, 0, 0);
        }
    }
}

cudaError_t launch_nms(const PipeP& P, int nframes, const int* d_integral, const float* d_resp, sb_point* d_points,
                       int* d_counts, unsigned* d_cand, int* d_cand_count, int cand_cap, unsigned* d_done, int* d_work,
                       int* d_work_orient, int* d_cls_cnt, cudaStream_t st) {
    // standard octaves (5 layers, the two cell lattices at most 2 samples apart): shared-memory scan, one CTA per tile
    cudaError_t e = cudaSuccess;
    bool tiled = P.max_scale == 5;
    int ctiles = 0;
    for (int o = 0; o < P.noctaves; o++) {
        const OctaveP& q = P.oct[o];
        if (q.nmb != 2 || abs(q.mb[0] - q.mb[1]) > 2) tiled = false;
        ctiles += q.nms_tx * q.nms_ty;
    }
    if (tiled)
        e = launch_dep(nms_scan_tile_kernel, dim3(ctiles, nframes), dim3(32, 8), 0, st, P, d_resp, d_cand, d_cand_count, cand_cap);
    else
        e = launch_dep(nms_scan_kernel, dim3(P.nms_tiles, nframes), dim3(32, 8), 0, st, P, d_resp, d_cand, d_cand_count, cand_cap);
    // ~5 k candidates per 1080p frame: 48 CTAs of 128 threads cover them in one pass, more are looped over
    if (e != cudaSuccess) return e;
    return launch_dep(nms_refine_kernel, dim3(48, nframes), dim3(128), 0, st, P, d_integral, d_resp, d_cand, d_cand_count, cand_cap, d_points, d_counts,
                      d_done, d_work, d_work_orient, d_cls_cnt);
}

}  // namespace sb
