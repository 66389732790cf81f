// match.cu -- brute-force descriptor matching: S = F1 * F2^T, per-row best / second / index, and
// the five match fields of SurfPoint.
//
// Replaces cuFindMaxCorr / findMaxCorr (surfd.cu:2535-2671, 3550-3566). Reference semantics kept:
//   * score = dot product accumulated as one FFMA chain over d = 0..nf-1 (surfd.cu:2591-2609);
//   * candidates are the first n2 - n2%32 descriptors of set 2 (surfd.cu:2569);
//   * candidates are partitioned into 8 groups g = (p2 % 32) / 4, each keeping a running top-2
//     with strict > from (0, 0, -1) in increasing p2 (surfd.cu:2610-2625);
//   * the merge starts from group 0 and sees only the maxima of the other groups
//     (surfd.cu:2646-2664); ambiguity = second / (best + 1e-6).
// Differences by design: writes are bounded to n1 rows (the reference writes whole 32-row blocks)
// and match_x/y are left 0 when nothing matched (the reference reads surf2[-1]).
//
// Round-1 kernel: exact fp32 CUDA-core version (row of F1 cached in registers, F2 tiles broadcast
// from shared memory). The tcgen05 tensor-core version replaces it once detect+describe is pinned.
#include "common.cuh"

namespace sb {

template <int NF>
__global__ void __launch_bounds__(256)
match_kernel(sb_point* __restrict__ pts1, int n1, const float* __restrict__ f1, const sb_point* __restrict__ pts2,
             int n2, const float* __restrict__ f2) {
    __shared__ __align__(16) float tile2[32][NF];
    __shared__ float s_max[8][32], s_sec[8][32];
    __shared__ int s_idx[8][32];
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int p1 = blockIdx.x * 32 + lane;
    const int p1c = min(p1, n1 - 1);
    float a[NF];
    {
        const float4* src = reinterpret_cast<const float4*>(f1 + (size_t)p1c * NF);
#pragma unroll
        for (int d = 0; d < NF / 4; d++) {
            const float4 t = __ldg(src + d);
            a[4 * d] = t.x; a[4 * d + 1] = t.y; a[4 * d + 2] = t.z; a[4 * d + 3] = t.w;
        }
    }
    float mx = 0.f, sc = 0.f;
    int id = -1;
    const int ncand = n2 - (n2 & 31);
    for (int bp2 = 0; bp2 < ncand; bp2 += 32) {
        __syncthreads();
        {
            const float4* src = reinterpret_cast<const float4*>(f2 + (size_t)bp2 * NF);
            float4* dst = reinterpret_cast<float4*>(&tile2[0][0]);
            for (int t = threadIdx.x; t < 32 * NF / 4; t += 256) dst[t] = __ldg(src + t);
        }
        __syncthreads();
#pragma unroll
        for (int dy = 0; dy < 4; dy++) {
            const float4* b = reinterpret_cast<const float4*>(&tile2[4 * g + dy][0]);
            float s = 0.f;
#pragma unroll
            for (int d = 0; d < NF / 4; d++) {
                const float4 t = b[d];
                s = __fmaf_rn(a[4 * d], t.x, s);
                s = __fmaf_rn(a[4 * d + 1], t.y, s);
                s = __fmaf_rn(a[4 * d + 2], t.z, s);
                s = __fmaf_rn(a[4 * d + 3], t.w, s);
            }
            if (s > mx) { sc = mx; mx = s; id = bp2 + 4 * g + dy; }
            else if (s > sc) sc = s;
        }
    }
    s_max[g][lane] = mx; s_sec[g][lane] = sc; s_idx[g][lane] = id;
    __syncthreads();
    if (g == 0 && p1 < n1) {
        float m = s_max[0][lane], s2 = s_sec[0][lane];
        int idx = s_idx[0][lane];
#pragma unroll
        for (int y = 0; y < 8; y++) {
            const int iy = s_idx[y][lane];
            const float my = s_max[y][lane];
            if (idx != iy) {
                if (my > m) { s2 = fmaxf(m, s2); m = my; idx = iy; }
                else if (my > s2) s2 = my;
            }
        }
        sb_point* p = pts1 + p1;
        p->score = m;
        p->match = idx;
        p->match_x = idx >= 0 ? pts2[idx].x : 0.f;
        p->match_y = idx >= 0 ? pts2[idx].y : 0.f;
        p->ambiguity = __fdiv_rn(s2, __fadd_rn(m, 1e-6f));
    }
}

cudaError_t launch_match(sb_point* d_pts1, int n1, const float* d_f1, const sb_point* d_pts2, int n2, const float* d_f2,
                         int nfeatures, cudaStream_t st) {
    if (n1 <= 0) return cudaSuccess;
    const dim3 grid((n1 + 31) / 32), block(256);
    if (nfeatures == 64) match_kernel<64><<<grid, block, 0, st>>>(d_pts1, n1, d_f1, d_pts2, n2, d_f2);
    else if (nfeatures == 128) match_kernel<128><<<grid, block, 0, st>>>(d_pts1, n1, d_f1, d_pts2, n2, d_f2);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

}  // namespace sb
