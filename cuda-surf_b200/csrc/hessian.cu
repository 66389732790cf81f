// hessian.cu -- box-filter determinant-of-Hessian for every octave and layer of a frame batch in
// ONE launch.
//
// Replaces the per-octave sequence of the reference: 2x cuHalfImage + cuCalcHessianMulti with three
// synchronous cudaMemcpyToSymbol each (surf.cpp:248-294, surfd.cu:321-331, 445-481, 2829-2894).
// A response value depends only on (integral, lobe, centre), so no octave waits for another:
//   * blockIdx.x enumerates 32x8 output tiles over all octaves (tile table in PipeP),
//     blockIdx.y is the frame;
//   * a thread evaluates all computed layers of its octave at one sample, so the rows it
//     gathers stay hot in L1 across layers;
//   * layers max_scale-3 / max_scale-1 are written through to layers 0 / 1 of the next octave at
//     even samples (what halfImage copies), chained for parameter sets where a copied layer is
//     itself a source.
// Outside [b1, dim-b1) a layer is never written and stays zero from allocation, which the
// reference's NMS relies on.
#include "common.cuh"

namespace sb {

// det = r^2 (Dxx*Dyy - (0.6*Dxy)^2) * norm with the exact operation order of the reference's
// sm_100a SASS (surfd.cu:353-366): t=(float(Dxy)*0.6f)^2; det=fma(Dxx,Dyy,-t); det*=r*r; det*=norm
__device__ __forceinline__ float hessian_response(const int* __restrict__ I, int ip, int cx, int cy, int l, float norm) {
    const int x2 = l >> 1, x3 = x2 + x2, x4 = x2 + x3;
    // Dxx: (2l+2*x2+1) x (2*x3+1) box minus 3x its central (2*x2+1)-wide part; rows shared
    int dxx, dyy, dxy;
    {
        const int* r0 = I + (cy - x3) * ip;
        const int* r1 = I + (cy + x3 + 1) * ip;
        const int a0 = cx - l - x2, a1 = cx - x2, a2 = cx + x2 + 1, a3 = cx + l + x2 + 1;
        const int wide = __ldg(r1 + a3) + __ldg(r0 + a0) - __ldg(r0 + a3) - __ldg(r1 + a0);
        const int mid = __ldg(r1 + a2) + __ldg(r0 + a1) - __ldg(r0 + a2) - __ldg(r1 + a1);
        dxx = wide - 3 * mid;
    }
    {
        const int c0 = cx - x3, c1 = cx + x3 + 1;
        const int* r0 = I + (cy - l - x2) * ip;
        const int* r1 = I + (cy - x2) * ip;
        const int* r2 = I + (cy + x2 + 1) * ip;
        const int* r3 = I + (cy + l + x2 + 1) * ip;
        const int tall = __ldg(r3 + c1) + __ldg(r0 + c0) - __ldg(r0 + c1) - __ldg(r3 + c0);
        const int mid = __ldg(r2 + c1) + __ldg(r1 + c0) - __ldg(r1 + c1) - __ldg(r2 + c0);
        dyy = tall - 3 * mid;
    }
    {
        // four (x4+1)^2 quadrant boxes sharing the centre pixel row/column
        const int* ra = I + (cy - x4) * ip;
        const int* rb = I + cy * ip;
        const int* rc = I + (cy + 1) * ip;
        const int* rd = I + (cy + x4 + 1) * ip;
        const int A = cx - x4, B = cx, C = cx + 1, D = cx + x4 + 1;
        const int tr = __ldg(rc + D) + __ldg(ra + B) - __ldg(ra + D) - __ldg(rc + B);
        const int bl = __ldg(rd + C) + __ldg(rb + A) - __ldg(rb + C) - __ldg(rd + A);
        const int br = __ldg(rd + D) + __ldg(rb + B) - __ldg(rb + D) - __ldg(rd + B);
        const int tl = __ldg(rc + C) + __ldg(ra + A) - __ldg(ra + C) - __ldg(rc + A);
        dxy = tr + bl - br - tl;
    }
    const float fxy = __fmul_rn(0.6f, __int2float_rn(dxy));
    const float t = __fmul_rn(fxy, fxy);
    float det = __fmaf_rn(__int2float_rn(dxx), __int2float_rn(dyy), -t);
    constexpr float r255 = 0.003921568627f;
    constexpr float rr = r255 * r255;  // float product, folded at compile time like the reference's r*r
    det = __fmul_rn(det, rr);
    return __fmul_rn(det, norm);
}

__global__ void __launch_bounds__(256)
hessian_kernel(const __grid_constant__ PipeP P, const int* __restrict__ Ibase, float* __restrict__ Rbase) {
    const int f = blockIdx.y;
    const int tile = blockIdx.x;
    int o = 0;
    while (o + 1 < P.noctaves && tile >= P.oct[o + 1].hess_tile0) o++;
    const OctaveP& q = P.oct[o];
    const int lt = tile - q.hess_tile0;
    const int ty = lt / q.hess_tx, tx = lt - ty * q.hess_tx;
    const int ix = tx * 32 + threadIdx.x, iy = ty * 8 + threadIdx.y;
    if (ix >= q.sw || iy >= q.sh) return;

    const int* I = Ibase + (size_t)f * P.istride + P.ip;
    float* Rf = Rbase + (size_t)f * P.rstride;
    const int cx = q.delta * ix, cy = q.delta * iy;
    const int ms = P.max_scale;
    for (int i = 0; i < q.nl; i++) {
        const int b = q.b1[i];
        if (ix < b || ix >= q.sw - b || iy < b || iy >= q.sh - b) continue;
        const float v = hessian_response(I, P.ip, cx, cy, q.l[i], q.norm[i]);
        int s = q.s0 + i;
        Rf[q.resp_off + (size_t)s * q.osz + (size_t)iy * q.sp + ix] = v;
        // write-through of what halfImage would copy into the next octave(s)
        int oo = o, x = ix, y = iy;
        while (oo + 1 < P.noctaves && (s == ms - 3 || s == ms - 1) && (((x | y) & 1) == 0)) {
            const OctaveP& n = P.oct[oo + 1];
            x >>= 1; y >>= 1;
            if (x >= n.sw || y >= n.sh) break;
            s = (s == ms - 3) ? 0 : 1;
            oo++;
            Rf[n.resp_off + (size_t)s * n.osz + (size_t)y * n.sp + x] = v;
        }
    }
}

cudaError_t launch_hessian(const PipeP& P, int nframes, const int* d_integral, float* d_resp, cudaStream_t st) {
    const dim3 grid(P.hess_tiles, nframes), block(32, 8);
    hessian_kernel<<<grid, block, 0, st>>>(P, d_integral, d_resp);
    return cudaGetLastError();
}

}  // namespace sb
