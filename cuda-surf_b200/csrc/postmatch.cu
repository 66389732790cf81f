// postmatch.cu -- consumer-side post-processing of Surfor::match results.
//
// The reference applies no acceptance test: main.cpp:59-70 draws every row's best candidate, and the "0.8 ratio test"
// of BASELINE config 5 is a filter `ambiguity < 0.8` the consumer runs on the SurfPoint fields that findMaxCorr wrote
// (surfd.cu:2665-2669). This kernel does that filter on the device and compacts the survivors into (idx1, idx2)
// pairs, optionally also requiring equal Laplacian signs (SurfPoint::laplace exists for this and is unused by the
// reference matcher) and a symmetric cross-check (set 2 matched back to set 1 by a second sb_match call).
// One CTA, block-wide scan per 1024-row chunk: the output order is the row order, so results are deterministic.
#include "common.cuh"

namespace sb {

__global__ void __launch_bounds__(1024)
match_filter_kernel(const sb_point* __restrict__ p1, int n1, const sb_point* __restrict__ p2, int n2, float max_ambiguity,
                    int flags, sb_pair* __restrict__ out, int cap, int* __restrict__ count) {
    __shared__ int wsum[32];
    __shared__ int base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n1; i0 += 1024) {
        const int i = i0 + tid;
        bool keep = false;
        int j = -1;
        float score = 0.f, amb = 0.f;
        if (i < n1) {
            const sb_point a = p1[i];
            j = a.match; score = a.score; amb = a.ambiguity;
            keep = j >= 0 && j < n2 && amb < max_ambiguity;
            if (keep && (flags & (SB_FILTER_LAPLACE | SB_FILTER_CROSS))) {
                const sb_point b = p2[j];
                if ((flags & SB_FILTER_LAPLACE) && b.laplace != a.laplace) keep = false;
                if ((flags & SB_FILTER_CROSS) && b.match != i) keep = false;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) wsum[warp] = __popc(m);
        __syncthreads();
        int off = base;
        for (int w = 0; w < warp; w++) off += wsum[w];
        if (keep) {
            const int slot = off + __popc(m & ((1u << lane) - 1u));
            if (slot < cap) { sb_pair pr; pr.idx1 = i; pr.idx2 = j; pr.score = score; pr.ambiguity = amb; out[slot] = pr; }
        }
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < 32; w++) t += wsum[w]; base += t; }
        __syncthreads();
    }
    if (tid == 0) *count = min(base, cap);
}

cudaError_t launch_match_filter(const sb_point* d_pts1, int n1, const sb_point* d_pts2, int n2, float max_ambiguity, int flags,
                                sb_pair* d_pairs, int cap, int* d_count, cudaStream_t st) {
    match_filter_kernel<<<1, 1024, 0, st>>>(d_pts1, n1, d_pts2, n2, max_ambiguity, flags, d_pairs, cap, d_count);
    return cudaGetLastError();
}

}  // namespace sb
