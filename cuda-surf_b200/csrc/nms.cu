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

// grid (nms_tiles, nframes), block 32x8: one thread per 2x2(xy) x 2(scale) cell.
__global__ void __launch_bounds__(256)
nms_kernel(const __grid_constant__ PipeP P, const int* __restrict__ Ibase, const float* __restrict__ Rbase,
           sb_point* __restrict__ points, int* __restrict__ counts) {
    const int f = blockIdx.y;
    const int tile = blockIdx.x;
    int o = 0;
    while (o + 1 < P.noctaves && tile >= P.oct[o + 1].nms_tile0) o++;
    const OctaveP& q = P.oct[o];
    int lt = tile - q.nms_tile0;
    const int per_z = q.nms_tx * q.nms_ty;
    const int z = lt / per_z;
    lt -= z * per_z;
    const int ty = lt / q.nms_tx, tx = lt - ty * q.nms_tx;
    const int lane = threadIdx.x;
    const int xc = tx * 32 + lane, yc = ty * 8 + threadIdx.y;

    const int ms = P.max_scale;
    const int k = 2 * z + 1;
    const int mb = q.mb[z];
    const int i = mb + 2 * yc, j = mb + 2 * xc;
    const int sw = q.sw, sh = q.sh, sp = q.sp, osz = q.osz;
    const float* src = Rbase + (size_t)f * P.rstride + q.resp_off;

    bool keep = false;
    float kx = 0.f, ky = 0.f, kscale = 0.f, kstrength = 0.f;
    int klap = 1;

    if (k < ms - 1 && i < sh - mb && j < sw - mb) {
        const float* c0 = src + (size_t)k * osz + (size_t)i * sp + j;
        const float* c1 = c0 + osz;
        // cell maximum in the reference's scan order, strict >
        float best = __ldg(c0);
        int cas = 0;
        float v;
        v = __ldg(c0 + 1);      if (v > best) { best = v; cas = 1; }
        v = __ldg(c0 + sp);     if (v > best) { best = v; cas = 2; }
        v = __ldg(c0 + sp + 1); if (v > best) { best = v; cas = 3; }
        v = __ldg(c1);          if (v > best) { best = v; cas = 4; }
        v = __ldg(c1 + 1);      if (v > best) { best = v; cas = 5; }
        v = __ldg(c1 + sp);     if (v > best) { best = v; cas = 6; }
        v = __ldg(c1 + sp + 1); if (v > best) { best = v; cas = 7; }
        bool cand = !(best < __fmul_rn(P.thresh, 0.8f)) && !(k + 1 == ms - 1 && cas > 3);
        int s = k + (cas >> 2), r = i + ((cas >> 1) & 1), c = j + (cas & 1);
        if (cand) {
            // outward directions: the cell's other member along each axis sits at -d
            const int ds = (cas & 4) ? 1 : -1, dr = (cas & 2) ? 1 : -1, dc = (cas & 1) ? 1 : -1;
            const float* ctr = src + (size_t)s * osz + (size_t)r * sp + c;
            // outer layer s+ds: all nine
            const float* L = ctr + ds * osz;
#pragma unroll
            for (int a = -1; a <= 1; a++)
#pragma unroll
                for (int b = -1; b <= 1; b++)
                    if (best < __ldg(L + a * sp + b)) cand = false;
            // own layer s and inner layer s-ds: the five positions outside the cell's 2x2 footprint
#pragma unroll
            for (int li = 0; li < 2; li++) {
                const float* M = ctr - li * ds * osz;
                const float* rowo = M + dr * sp;  // outward row: three
                if (best < __ldg(rowo - 1)) cand = false;
                if (best < __ldg(rowo)) cand = false;
                if (best < __ldg(rowo + 1)) cand = false;
                if (best < __ldg(M + dc)) cand = false;            // (r, c+dc)
                if (best < __ldg(M - dr * sp + dc)) cand = false;  // (r-dr, c+dc)
            }
        }
        if (cand) {
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

__global__ void clamp_counts_kernel(int* counts, int n, int max_pts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) counts[i] = min(counts[i], max_pts);
}

cudaError_t launch_nms(const PipeP& P, int nframes, const int* d_integral, const float* d_resp, sb_point* d_points,
                       int* d_counts, cudaStream_t st) {
    const dim3 grid(P.nms_tiles, nframes), block(32, 8);
    nms_kernel<<<grid, block, 0, st>>>(P, d_integral, d_resp, d_points, d_counts);
    return cudaGetLastError();
}

cudaError_t launch_clamp_counts(int* d_counts, int nframes, int max_pts, cudaStream_t st) {
    clamp_counts_kernel<<<(nframes + 255) / 256, 256, 0, st>>>(d_counts, nframes, max_pts);
    return cudaGetLastError();
}

}  // namespace sb
