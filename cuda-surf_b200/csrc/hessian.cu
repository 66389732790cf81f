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

// det = r^2 (Dxx*Dyy - (0.6*Dxy)^2) * norm with the exact operation order of the reference's sm_100a SASS
// (surfd.cu:353-366): t=(float(Dxy)*0.6f)^2; det=fma(Dxx,Dyy,-t); det*=r*r; det*=norm
__device__ __forceinline__ float hessian_det(int dxx, int dyy, int dxy, float norm) {
    const float fxy = __fmul_rn(0.6f, __int2float_rn(dxy));
    const float t = __fmul_rn(fxy, fxy);
    float det = __fmaf_rn(__int2float_rn(dxx), __int2float_rn(dyy), -t);
    constexpr float r255 = 0.003921568627f;
    constexpr float rr = r255 * r255;  // float product, folded at compile time like the reference's r*r
    det = __fmul_rn(det, rr);
    return __fmul_rn(det, norm);
}

// The 32 box corners of a sample lie on 10 rows and 10 columns (same corners as getSum, surfd.cu:334-343):
//   rows    Dxx: -x3, x3+1 | Dyy: -l-x2, -x2, x2+1, l+x2+1 | Dxy: -x4, 0, 1, x4+1
//   columns Dxx: -l-x2, -x2, x2+1, l+x2+1 | Dyy: -x3, x3+1 | Dxy: -x4, 0, 1, x4+1
// A thread keeps its lobe and column for all of its samples, so the 10 row offsets (elements) and the 10 column
// positions -- in the column-phase layout of the second integral copy (common.cuh: phase_col) -- are computed once;
// a sample then costs 10 row pointers (IMAD.WIDE, made opaque so they are not re-derived per load) and one
// IMAD.WIDE + LDG per corner. Written naively the compiler spent 7 integer instructions of 64-bit address
// arithmetic per load (ncu: 340 instructions per sample).
struct CornerGeom { int roff[10]; int col[10]; };

__device__ __forceinline__ CornerGeom corner_geom(int l, int cx, int ip) {
    const int x2 = l >> 1, x3 = x2 + x2, x4 = x2 + x3;
    const int dy[10] = {-x3, x3 + 1, -l - x2, -x2, x2 + 1, l + x2 + 1, -x4, 0, 1, x4 + 1};
    const int dx[10] = {-l - x2, -x2, x2 + 1, l + x2 + 1, -x3, x3 + 1, -x4, 0, 1, x4 + 1};
    CornerGeom g;
#pragma unroll
    for (int k = 0; k < 10; k++) { g.roff[k] = dy[k] * ip; g.col[k] = phase_col(cx + dx[k], ip); }
    return g;
}

// `row`: the sample's centre row in the column-phase integral copy
__device__ __forceinline__ float hessian_response(const int* __restrict__ row, const CornerGeom& g, float norm) {
    const int* r[10];
    int c[10];
#pragma unroll
    for (int k = 0; k < 10; k++) {
        // opaque 32-bit offsets: otherwise they are widened once outside the loop and every address is a 64-bit add pair
        int ro = g.roff[k];
        c[k] = g.col[k];
        asm volatile("" : "+r"(ro), "+r"(c[k]));
        r[k] = row + ro;
        asm volatile("" : "+l"(r[k]));
    }
#define G_(ri, ci) __ldg(r[ri] + c[ci])
    const int wide = G_(1, 3) + G_(0, 0) - G_(0, 3) - G_(1, 0);
    const int midx = G_(1, 2) + G_(0, 1) - G_(0, 2) - G_(1, 1);
    const int dxx = wide - 3 * midx;
    const int tall = G_(5, 5) + G_(2, 4) - G_(2, 5) - G_(5, 4);
    const int midy = G_(4, 5) + G_(3, 4) - G_(3, 5) - G_(4, 4);
    const int dyy = tall - 3 * midy;
    // four (x4+1)^2 quadrant boxes sharing the centre pixel row/column: rows a=6, b=7, c=8, d=9; columns A=6, B=7, C=8, D=9
    const int tr = G_(8, 9) + G_(6, 7) - G_(6, 9) - G_(8, 7);
    const int bl = G_(9, 8) + G_(7, 6) - G_(7, 8) - G_(9, 6);
    const int br = G_(9, 9) + G_(7, 7) - G_(7, 9) - G_(9, 7);
    const int tl = G_(8, 8) + G_(6, 6) - G_(6, 8) - G_(8, 6);
    const int dxy = tr + bl - br - tl;
#undef G_
    return hessian_det(dxx, dyy, dxy, norm);
}

// grid (tiles * layers, nframes), block 32x8; the layer is the fastest-varying part of blockIdx.x, so the layers of a
// tile -- and all tiles of a frame -- run while that frame's integral is in L2 (with the layer as blockIdx.z every layer
// swept all 64 frames again: ncu, 25 MB of DRAM reads per frame for the 8.9 MB integral). A CTA covers 32 x kHessRows outputs of one layer (a thread: kHessRows/8 rows,
// 8 apart): with a 32x8 tile every CTA pulled its whole 70-pixel filter halo through L2 for 256 outputs (ncu: 94 MB of
// L2->L1 traffic per frame for the 8.9 MB integral); the taller tile reuses the halo from L1. (Looping the layers inside
// the CTA as well was 1 % faster in a 64-frame batch and 30 % slower for a single frame.)
__global__ void __launch_bounds__(256, 4)
hessian_kernel(const __grid_constant__ PipeP P, const int* __restrict__ Iphase, float* __restrict__ Rbase, int first_tile,
               int nlayers) {
    pdl_wait();
    const int f = blockIdx.y;
    // nlayers > 0: one CTA per (tile, layer), the layer fastest; nlayers == 0: one CTA per tile loops over its layers, which
    // share the tile's patch of the integral in L1 (used for batches, where there are CTAs enough without the split)
    const int tile = (nlayers > 0 ? blockIdx.x / nlayers : blockIdx.x) + first_tile;
    int o = 0;
    while (o + 1 < P.noctaves && tile >= P.oct[o + 1].hess_tile0) o++;
    const OctaveP& q = P.oct[o];
    const int i0 = nlayers > 0 ? blockIdx.x - (blockIdx.x / nlayers) * nlayers : 0;
    const int i1 = nlayers > 0 ? min(i0 + 1, q.nl) : q.nl;
    const int lt = tile - q.hess_tile0;
    const int ty = div_small(lt, q.inv_hess_tx), tx = lt - ty * q.hess_tx;
    const int ix = tx * 32 + threadIdx.x;
    if (ix >= q.sw) return;
    const int* I = Iphase + (size_t)f * P.istride + P.ip;
    float* Rf = Rbase + (size_t)f * P.rstride;
    const int cx = q.delta * ix;
    const int ms = P.max_scale;
#pragma unroll 1
    for (int i = i0; i < i1; i++) {
        const int b = q.b1[i];
        if (ix < b || ix >= q.sw - b) continue;
        const CornerGeom g = corner_geom(q.l[i], cx, P.ip);
        const float norm = q.norm[i];
        const int rowstep = q.delta * P.ip;
#pragma unroll 1
        for (int u = 0; u < kHessRows / 8; u++) {
            const int iy = ty * kHessRows + threadIdx.y + 8 * u;
            if (iy < b || iy >= q.sh - b) continue;
            const float v = hessian_response(I + (size_t)iy * rowstep, g, norm);
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
}


// ------------------------------------------------------------------ octave 0 from shared memory
//
// Octave 0 holds 84 % of all response samples and its filters are small (9..33 px), so a tile of
// outputs can share one staged patch of the integral image. The gather kernel above is bound by
// the L1 data stage: neighbouring outputs are 2 px apart, so a warp's 4-byte gathers use half of
// every 32-byte sector (ncu: 77 % l1tex throughput, 10 sectors per request). Here a CTA stages the
// 96x64 patch under a 32x16 output tile once, DE-INTERLEAVED by row/column parity into four planes,
// so the 32 lanes of a warp read 32 consecutive words for every corner: one conflict-free
// shared-memory wavefront per gather instead of 2.5 L1 cycles, and every corner address is
// (thread base + compile-time constant) because the lobe is a template parameter.
// Only for the reference's default geometry (sampling 2, lobes 3,5,7,9,11); anything else takes the
// generic kernel.
constexpr int kTW = 32, kTH = 16;            // outputs per tile
constexpr int kHalo = 16;                    // patch origin = D * tile origin - kHalo (the widest corner offset is 17, the last sample D px short of the tile's end)
// D = sampling step of octave 0 in pixels: 2 (the reference's default) or 4 (doubled=true, where the 2x frame is sampled every
// 4th pixel). The staged patch is D*32+32 x D*16+32 integral elements, de-interleaved into D x D (row phase, column
// phase) planes, so the 32 lanes of a warp -- D pixels apart -- read 32 consecutive words for every corner.
template <int D> struct O0 {
    static constexpr int PW = D * kTW + 2 * kHalo, PH = D * kTH + 2 * kHalo;  // 96 x 64 (D = 2), 160 x 96 (D = 4)
    static constexpr int QW = PW / D, QH = PH / D;                              // plane dims
    static constexpr int Plane = QW * QH;
    // D = 2: the two planes of odd patch rows start 16 words later: a staging warp covers the end of one patch row and the
    // start of the next (24 int4 per row), whose 64-bit stores otherwise meet in banks 0..15 (ncu: 7 % of the wavefronts
    // were conflicts)
    static constexpr int RowPhase = D * Plane + (D == 2 ? 16 : 0);
    static constexpr int Words = D * RowPhase;
    static constexpr int Int4PerRow = PW / 4;
    static constexpr int Int4 = PH * Int4PerRow;  // 1536 (D = 2), 3840 (D = 4)
    static_assert(Int4 % 256 == 0, "staging assumes a whole number of int4 per thread");
};

// word offset of patch element (cy + dy, cx + dx) relative to the thread base (ly*QW + lx)
template <int D>
__host__ __device__ constexpr int corner_off(int dx, int dy) {
    return ((dy + kHalo) & (D - 1)) * O0<D>::RowPhase + ((dx + kHalo) & (D - 1)) * O0<D>::Plane + ((dy + kHalo) / D) * O0<D>::QW +
           ((dx + kHalo) / D);
}

// ctr = the four corners (0,0), (1,0), (0,1), (1,1) [as (dx,dy)] around the sample: every layer's Dxy uses them, so
// they are read once per sample instead of once per layer (16 of the 160 shared-memory words of a sample)
template <int L, int D>
__device__ __forceinline__ float response_smem(const int* __restrict__ b, float norm, const int (&ctr)[4]) {
    constexpr int x2 = L / 2, x3 = 2 * x2, x4 = 3 * x2;
#define C_(dx, dy) b[corner_off<D>(dx, dy)]
    // same corners as hessian_response(): rows/cols are those of getSum (surfd.cu:334-343)
    const int wide = C_(L + x2 + 1, x3 + 1) + C_(-L - x2, -x3) - C_(L + x2 + 1, -x3) - C_(-L - x2, x3 + 1);
    const int midx = C_(x2 + 1, x3 + 1) + C_(-x2, -x3) - C_(x2 + 1, -x3) - C_(-x2, x3 + 1);
    const int dxx = wide - 3 * midx;
    const int tall = C_(x3 + 1, L + x2 + 1) + C_(-x3, -L - x2) - C_(x3 + 1, -L - x2) - C_(-x3, L + x2 + 1);
    const int midy = C_(x3 + 1, x2 + 1) + C_(-x3, -x2) - C_(x3 + 1, -x2) - C_(-x3, x2 + 1);
    const int dyy = tall - 3 * midy;
    const int tr = C_(x4 + 1, 1) + C_(0, -x4) - C_(x4 + 1, -x4) - ctr[2];
    const int bl = C_(1, x4 + 1) + C_(-x4, 0) - ctr[1] - C_(-x4, x4 + 1);
    const int br = C_(x4 + 1, x4 + 1) + ctr[0] - C_(x4 + 1, 0) - C_(0, x4 + 1);
    const int tl = ctr[3] + C_(-x4, -x4) - C_(1, -x4) - C_(-x4, 1);
    const int dxy = tr + bl - br - tl;
#undef C_
    const float fxy = __fmul_rn(0.6f, __int2float_rn(dxy));
    const float t = __fmul_rn(fxy, fxy);
    float det = __fmaf_rn(__int2float_rn(dxx), __int2float_rn(dyy), -t);
    constexpr float r255 = 0.003921568627f;
    constexpr float rr = r255 * r255;
    det = __fmul_rn(det, rr);
    return __fmul_rn(det, norm);
}

template <int L, int LAYER, int D>
__device__ __forceinline__ void layer_smem(const PipeP& P, const int* __restrict__ b, float* __restrict__ Rf, int ix, int iy,
                                           const int (&ctr)[4]) {
    const OctaveP& q = P.oct[0];
    const int bd = q.b1[LAYER];
    if (ix < bd || ix >= q.sw - bd || iy < bd || iy >= q.sh - bd) return;
    const float v = response_smem<L, D>(b, q.norm[LAYER], ctr);
    Rf[q.resp_off + (size_t)LAYER * q.osz + (size_t)iy * q.sp + ix] = v;
    // layers 2 and 4 are what halfImage copies into layers 0 and 1 of octave 1 (surf.cpp:250-258)
    if ((LAYER == 2 || LAYER == 4) && P.noctaves > 1 && ((ix | iy) & 1) == 0) {
        const OctaveP& n = P.oct[1];
        const int x = ix >> 1, y = iy >> 1;
        if (x < n.sw && y < n.sh) Rf[n.resp_off + (size_t)(LAYER == 2 ? 0 : 1) * n.osz + (size_t)y * n.sp + x] = v;
    }
}

// grid (tiles_x * tiles_y, nframes), 256 threads, dynamic shared memory O0<D>::Words ints (24.6 KB for D = 2, 61.4 KB for D = 4)
template <int D>
__global__ void __launch_bounds__(256, D == 2 ? 4 : 3)
hessian_o0_kernel(const __grid_constant__ PipeP P, const int* __restrict__ Ibase, float* __restrict__ Rbase, int tiles_x) {
    pdl_wait();
    using G = O0<D>;
    extern __shared__ __align__(16) int patch[];
    const int f = blockIdx.y;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int* I = Ibase + (size_t)f * P.istride + P.ip;
    const int X0 = D * kTW * tx - kHalo, Y0 = D * kTH * ty - kHalo;
    // stage: PH rows x PW/4 int4; element x of a row goes to column-phase plane x mod D
    // (all loads of a thread are in flight before the first store waits on one: as a plain loop every position was its own
    // memory round trip)
    constexpr int PER = G::Int4 / 256;  // 6 (D = 2), 15 (D = 4)
    constexpr int NB = PER % 6 == 0 ? 6 : 5;  // loads in flight per thread: all 6, or 3 batches of 5
#pragma unroll
    for (int it0 = 0; it0 < PER; it0 += NB) {
        int4 v[NB];
#pragma unroll
        for (int u = 0; u < NB; u++) {
            const int it = it0 + u;
            if (it < PER) {
                const int t = threadIdx.x + 256 * it;
                const int row = t / G::Int4PerRow, k = t - row * G::Int4PerRow;
                const int y = Y0 + row, x = X0 + 4 * k;
                const bool in = y >= 0 && y < P.ih && x >= 0 && x + 3 < P.ip;
                const int4 ld = __ldg(reinterpret_cast<const int4*>(I + (in ? (size_t)y * P.ip + x : (size_t)0)));
                v[u] = in ? ld : make_int4(0, 0, 0, 0);
            }
        }
#pragma unroll
        for (int u = 0; u < NB; u++) {
            const int it = it0 + u;
            if (it < PER) {
                const int t = threadIdx.x + 256 * it;
                const int row = t / G::Int4PerRow, k = t - row * G::Int4PerRow;
                int* dst = patch + (row & (D - 1)) * G::RowPhase + (row / D) * G::QW;
                if (D == 2) {
                    dst += 2 * k;
                    *reinterpret_cast<int2*>(dst) = make_int2(v[u].x, v[u].z);
                    *reinterpret_cast<int2*>(dst + G::Plane) = make_int2(v[u].y, v[u].w);
                } else {
                    dst += k;
                    dst[0] = v[u].x; dst[G::Plane] = v[u].y; dst[2 * G::Plane] = v[u].z; dst[3 * G::Plane] = v[u].w;
                }
            }
        }
    }
    __syncthreads();
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    float* Rf = Rbase + (size_t)f * P.rstride;
    const int ix = kTW * tx + lx;
#pragma unroll
    for (int hrow = 0; hrow < 2; hrow++) {
        const int lyy = ly + 8 * hrow;
        const int iy = kTH * ty + lyy;
        const int* b = patch + lyy * G::QW + lx;
        const int ctr[4] = {b[corner_off<D>(0, 0)], b[corner_off<D>(1, 0)], b[corner_off<D>(0, 1)], b[corner_off<D>(1, 1)]};
        layer_smem<3, 0, D>(P, b, Rf, ix, iy, ctr);
        layer_smem<5, 1, D>(P, b, Rf, ix, iy, ctr);
        layer_smem<7, 2, D>(P, b, Rf, ix, iy, ctr);
        layer_smem<9, 3, D>(P, b, Rf, ix, iy, ctr);
        layer_smem<11, 4, D>(P, b, Rf, ix, iy, ctr);
    }
}

cudaError_t launch_hessian(const PipeP& P, int nframes, const int* d_integral, const int* d_integral_ph, float* d_resp,
                           cudaStream_t st) {
    const OctaveP& q0 = P.oct[0];
    const bool fast0 = (P.sampling == 2 || P.sampling == 4) && P.init_lobe == 3 && P.max_scale == 5 && q0.s0 == 0 && q0.nl == 5;
    int first_tile = 0;
    if (fast0) {
        const int tiles_x = (q0.sw + kTW - 1) / kTW, tiles_y = (q0.sh + kTH - 1) / kTH;
        cudaError_t e;
        if (P.sampling == 2) {
            e = launch_dep(hessian_o0_kernel<2>, dim3(tiles_x * tiles_y, nframes), dim3(256), O0<2>::Words * sizeof(int), st, P, d_integral, d_resp, tiles_x);
        } else {
            e = cudaFuncSetAttribute(hessian_o0_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(O0<4>::Words * sizeof(int)));
            if (e == cudaSuccess)
                e = launch_dep(hessian_o0_kernel<4>, dim3(tiles_x * tiles_y, nframes), dim3(256), O0<4>::Words * sizeof(int), st, P, d_integral, d_resp, tiles_x);
        }
        if (e != cudaSuccess) return e;
        first_tile = P.noctaves > 1 ? P.oct[1].hess_tile0 : P.hess_tiles;
    }
    if (P.hess_tiles - first_tile > 0) {
        int maxnl = 0;
        for (int o = 0; o < P.noctaves; o++)
            if (P.oct[o].hess_tile0 >= first_tile && P.oct[o].nl > maxnl) maxnl = P.oct[o].nl;
        // (grid.y is the frame: the integral / response slots are indexed by blockIdx.y in every kernel)
        const bool fuse = nframes >= 16;
        const dim3 grid((P.hess_tiles - first_tile) * (fuse ? 1 : maxnl), nframes), block(32, 8);
        return launch_dep(hessian_kernel, grid, block, 0, st, P, d_integral_ph, d_resp, first_tile, fuse ? 0 : maxnl);
    }
    return cudaGetLastError();
}

}  // namespace sb
