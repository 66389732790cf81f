// integral.cu -- padded integral image, u8 -> int32, I[y+1][x+1] = sum_{j<=y,i<=x} img[j][i].
//
// Replaces cuIntegral = integralRow + integralCol (surfd.cu:129-165, 2683-2704): the reference runs
// one thread per image row, then one thread per column, each a serial loop over the other
// dimension. Here the image is cut into tiles of kBand rows x 256 output columns and the 2-D
// prefix is done reduce-then-scan:
//   pass A (integral_reduce): per tile, column sums T[band][X], row sums R[y][chunk] and the tile
//          total TT[band][chunk]                      -- reads the image once (HBM), writes ~4 %.
//   pass B (integral_scan):   per tile, carry row = look-back over T/TT of the bands above,
//          warp-shuffle row scans (8 px per lane from one 64-bit load), shared-memory column
//          scan, coalesced 128-byte row stores      -- re-reads the image from L2, writes the
//          integral exactly once.
// Compulsory traffic is img + integral; the reference moves the integral three times.
// Chunks are aligned in OUTPUT columns (X = x+1), so every integral row store is a full 128-byte
// line; the one-byte shift lands on the (cheap) image loads instead.
#include "common.cuh"

namespace sb {

constexpr int kBand = 32;     // rows per tile
constexpr int kChunk = 256;   // output columns per tile
constexpr int kThreads = 256;

// 8 pixels for output columns X = 256c + 8*lane + k  (image column X-1), zero outside the image: the raw load (8 bytes of
// this lane's aligned group + lane 0's byte from the previous chunk) and the unpack are separate, so a kernel can issue
// the loads of several rows before it waits for any of them.
struct RawRow { unsigned lo, hi, prev0; };

template <bool ALIGNED>
__device__ __forceinline__ RawRow load_raw(const uint8_t* __restrict__ row, int w, int c, int lane) {
    const int x0 = kChunk * c + 8 * lane;  // first image column of this lane's aligned 8-byte group
    RawRow q;
    q.lo = 0; q.hi = 0;
    if (ALIGNED && x0 + 7 < w) {
        const uint2 a = __ldg(reinterpret_cast<const uint2*>(row + x0));
        q.lo = a.x; q.hi = a.y;
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) if (x0 + k < w) q.lo |= (unsigned)__ldg(row + x0 + k) << (8 * k);
#pragma unroll
        for (int k = 0; k < 4; k++) if (x0 + 4 + k < w) q.hi |= (unsigned)__ldg(row + x0 + 4 + k) << (8 * k);
    }
    q.prev0 = (lane == 0 && x0 > 0) ? (unsigned)__ldg(row + x0 - 1) : 0u;
    return q;
}

__device__ __forceinline__ void unpack_shifted(const RawRow& q, int lane, int (&v)[8]) {
    unsigned prev = __shfl_up_sync(0xffffffffu, q.hi >> 24, 1);
    if (lane == 0) prev = q.prev0;
    const unsigned lo = q.lo, hi = q.hi;
    v[0] = (int)prev;  // one PRMT per byte
    v[1] = (int)__byte_perm(lo, 0, 0x4440); v[2] = (int)__byte_perm(lo, 0, 0x4441); v[3] = (int)__byte_perm(lo, 0, 0x4442);
    v[4] = (int)__byte_perm(lo, 0, 0x4443); v[5] = (int)__byte_perm(hi, 0, 0x4440); v[6] = (int)__byte_perm(hi, 0, 0x4441);
    v[7] = (int)__byte_perm(hi, 0, 0x4442);
}

__device__ __forceinline__ int warp_sum(int v) { return __reduce_add_sync(0xffffffffu, v); }  // one REDUX.SUM
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// Pass A. grid (nchunks, nbands, nframes), 256 threads.
template <bool ALIGNED>
__global__ void __launch_bounds__(kThreads)
integral_reduce(const __grid_constant__ PipeP P, const uint8_t* __restrict__ imgs, size_t image_stride, int pitch,
                int* __restrict__ T, int* __restrict__ R, int* __restrict__ TT, int* __restrict__ counts) {
    // the first kernel of a frame also zeroes the frame's keypoint counter (it was a memset node in front of every frame)
    if (counts && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) counts[blockIdx.z] = 0;
    __shared__ int part[8][kChunk];
    __shared__ int wtot[8];
    const int c = blockIdx.x, b = blockIdx.y, f = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint8_t* img = imgs + (size_t)f * image_stride;
    const int hpad = P.nbands * kBand;
    int acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = 0;
    // the four rows of this warp: all loads first (one exposed latency), then the sums
    RawRow raw[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int y = b * kBand + warp + 8 * u;
        if (y < P.h) raw[u] = load_raw<ALIGNED>(img + (size_t)y * pitch, P.w, c, lane);
        else { raw[u].lo = 0; raw[u].hi = 0; raw[u].prev0 = 0; }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int y = b * kBand + warp + 8 * u;
        int v[8];
        unpack_shifted(raw[u], lane, v);
        int s = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { s += v[k]; acc[k] += v[k]; }
        s = warp_sum(s);
        if (lane == 0) R[((size_t)f * hpad + y) * P.nchunks + c] = s;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) part[warp][8 * lane + k] = acc[k];
    __syncthreads();
    int col = 0;
#pragma unroll
    for (int wq = 0; wq < 8; wq++) col += part[wq][tid];
    T[(((size_t)f * P.nbands + b) * P.nchunks + c) * kChunk + tid] = col;
    const int ws = warp_sum(col);
    if (lane == 0) wtot[warp] = ws;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
#pragma unroll
        for (int wq = 0; wq < 8; wq++) t += wtot[wq];
        TT[((size_t)f * P.nbands + b) * P.nchunks + c] = t;
    }
}

// Pass B. grid (nchunks, nbands, nframes), 256 threads, 32 KB static shared memory.
// Columns first, rows last: the tile's pixels are summed DOWN the columns (starting from the column sums of the bands
// above), then every row is prefix-summed ALONG x by one warp with 8 consecutive output columns per lane. The values a
// lane ends with are final, so it stores them twice without any scatter: two 128-bit stores into the row-major image and,
// for k = 0..7, one word into plane k of the column-phase copy (common.cuh: phase_col) -- lane-consecutive, one full
// 128-byte line per plane. (With the column scan last, a warp held 32 consecutive columns of one row and the phase copy
// cost 8 store wavefronts per instruction: +2 us per 1080p frame.)
template <bool ALIGNED>
__global__ void __launch_bounds__(kThreads, 6)
integral_scan(const __grid_constant__ PipeP P, const uint8_t* __restrict__ imgs, size_t image_stride, int pitch,
              const int* __restrict__ T, const int* __restrict__ R, const int* __restrict__ TT, int* __restrict__ Iout,
              int* __restrict__ Iph) {
    pdl_wait();
    __shared__ __align__(16) int tile[kBand][kChunk];
    __shared__ int s_rowbase[kBand];
    __shared__ int s_offw[8];
    const int c = blockIdx.x, b = blockIdx.y, f = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nb = P.nbands, nc = P.nchunks;
    const uint8_t* img = imgs + (size_t)f * image_stride;
    const int hpad = nb * kBand;

    // Every global load of the tile is issued before the first use (one exposed memory latency per CTA, two barriers):
    // ---- raw pixels: warp per row, rows warp, warp+8, warp+16, warp+24
    RawRow raw[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int y = b * kBand + warp + 8 * u;
        if (y < P.h) raw[u] = load_raw<ALIGNED>(img + (size_t)y * pitch, P.w, c, lane);
        else { raw[u].lo = 0; raw[u].hi = 0; raw[u].prev0 = 0; }
    }
    // The three look-backs below are sums of independent loads; each is written as groups of predicated loads so that a
    // group is ONE exposed memory latency (as plain loops with a running sum, every iteration waited for its load: up to
    // 8 + 5 + 7 serial round trips per CTA at 1080p).
    // ---- column carry: sum of this output column over the bands above
    int colc = 0;
    {
        const int* Tp = T + (((size_t)f * nb) * nc + c) * kChunk + tid;
        const size_t bstride = (size_t)nc * kChunk;
#pragma unroll 1
        for (int b0 = 0; b0 < b; b0 += 16) {
            int a[16];
#pragma unroll
            for (int k = 0; k < 16; k++) a[k] = (b0 + k < b) ? __ldg(Tp + (size_t)(b0 + k) * bstride) : 0;
#pragma unroll
            for (int k = 0; k < 16; k += 4) colc += (a[k] + a[k + 1]) + (a[k + 2] + a[k + 3]);
        }
    }
    // ---- totals of the tiles above and to the left: warp per band (4 bands in flight), lane per chunk
    int off = 0;
#pragma unroll 1
    for (int bb = warp; bb < b; bb += 32) {
        int a[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            a[k] = 0;
            if (bb + 8 * k < b)
                for (int cc = lane; cc < c; cc += 32) a[k] += __ldg(TT + ((size_t)f * nb + bb + 8 * k) * nc + cc);
        }
        off += (a[0] + a[1]) + (a[2] + a[3]);
    }
    // ---- row sums of the chunks to the left (warp 0, lane = row of the band), 8 chunks in flight
    int rleft = 0;
    if (warp == 0) {
        const int* Rp = R + ((size_t)f * hpad + b * kBand + lane) * nc;
#pragma unroll 1
        for (int c0 = 0; c0 < c; c0 += 8) {
            int a[8];
#pragma unroll
            for (int k = 0; k < 8; k++) a[k] = (c0 + k < c) ? __ldg(Rp + c0 + k) : 0;
            rleft += ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
        }
    }
    off = warp_sum(off);
    if (lane == 0) s_offw[warp] = off;
#pragma unroll
    for (int u = 0; u < 4; u++) {
        int v[8];
        unpack_shifted(raw[u], lane, v);
        int4* dst = reinterpret_cast<int4*>(&tile[warp + 8 * u][8 * lane]);
        dst[0] = make_int4(v[0], v[1], v[2], v[3]);
        dst[1] = make_int4(v[4], v[5], v[6], v[7]);
    }
    __syncthreads();
    // ---- row bases: integral just left of this chunk at every row of the band = tiles above-left + the chunks' row sums
    // down to that row
    if (warp == 0) {
        int o = 0;
#pragma unroll
        for (int wq = 0; wq < 8; wq++) o += s_offw[wq];
        s_rowbase[lane] = o + warp_incl_scan(rleft, lane);
    }
    // ---- column sums down the tile, one output column per thread
    {
        int run = colc;
#pragma unroll 8
        for (int r = 0; r < kBand; r++) {
            run += tile[r][tid];
            tile[r][tid] = run;
        }
    }
    __syncthreads();

    // ---- row scans and stores: warp per row
    const int rows = min(kBand, P.h - b * kBand);
    const int X0 = kChunk * c + 8 * lane;
    const int ipitch = P.ip, pw = ipitch >> 3;
    for (int r = warp; r < rows; r += 8) {
        const int4* src = reinterpret_cast<const int4*>(&tile[r][8 * lane]);
        const int4 lo = src[0], hi = src[1];
        int v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
        for (int k = 1; k < 8; k++) v[k] += v[k - 1];
        const int inc = warp_incl_scan(v[7], lane);
        const int base = s_rowbase[r] + inc - v[7];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] += base;
        const size_t rowoff = (size_t)f * P.istride + (size_t)(b * kBand + r + 2) * ipitch;  // guard row + zero row
        int* out = Iout + rowoff + X0;
        int* outp = Iph + rowoff + (X0 >> 3);
        if (X0 + 7 < P.iw) {
            reinterpret_cast<int4*>(out)[0] = make_int4(v[0], v[1], v[2], v[3]);
            reinterpret_cast<int4*>(out)[1] = make_int4(v[4], v[5], v[6], v[7]);
#pragma unroll
            for (int k = 0; k < 8; k++) outp[k * pw] = v[k];
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (X0 + k < P.iw) { out[k] = v[k]; outp[k * pw] = v[k]; }
        }
    }
}

// ------------------------------------------------------------------ doubled=true: 2x up-sampling
//
// The reference builds the integral of a 2x bilinear up-sampling of the frame (integralDoubleRow0U2 + five scan
// kernels, surfd.cu:166-318, 2707-2772). The up-sampled pixels are defined by surfd.cu:182-206: even/even the source
// pixel, otherwise rn() of the mean of the 2 or 4 neighbours; the image the scans cover is (2w-2) x (2h-2). Here the
// 2x image is materialised as u8 (one thread per 4 output pixels, 32-bit stores) and goes through the same
// reduce-then-scan kernels as any frame.
__global__ void __launch_bounds__(256)
upsample2x_kernel(const uint8_t* __restrict__ src, size_t src_stride, int src_pitch, int w, int h, uint8_t* __restrict__ dst,
                  size_t dst_stride, int dst_pitch) {
    const int W2 = 2 * w - 2, H2 = 2 * h - 2;
    const int X = 4 * (blockIdx.x * blockDim.x + threadIdx.x), Y = blockIdx.y;
    if (X >= W2 || Y >= H2) return;
    const uint8_t* r0 = src + blockIdx.z * src_stride + (size_t)(Y >> 1) * src_pitch + (X >> 1);
    const uint8_t* r1 = (Y & 1) ? r0 + src_pitch : r0;  // odd rows average two source rows
    // source columns x, x+1, x+2 (x+2 only feeds output X+3, which exists iff X+3 < W2)
    const int a0 = r0[0] + r1[0], a1 = r0[1] + r1[1];
    const int a2 = (X + 3 < W2) ? r0[2] + r1[2] : a1;
    unsigned v[4];
    if (Y & 1) {
        v[0] = __float2int_rn(__int2float_rn(a0) * 0.5f);
        v[1] = __float2int_rn(__int2float_rn(a0 + a1) * 0.25f);
        v[2] = __float2int_rn(__int2float_rn(a1) * 0.5f);
        v[3] = __float2int_rn(__int2float_rn(a1 + a2) * 0.25f);
    } else {  // r1 == r0: a = 2 * pixel
        v[0] = a0 >> 1;
        v[1] = __float2int_rn(__int2float_rn((a0 + a1) >> 1) * 0.5f);
        v[2] = a1 >> 1;
        v[3] = __float2int_rn(__int2float_rn((a1 + a2) >> 1) * 0.5f);
    }
    uint8_t* d = dst + blockIdx.z * dst_stride + (size_t)Y * dst_pitch + X;
    if (X + 3 < W2) {
        *reinterpret_cast<unsigned*>(d) = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
    } else {
        for (int k = 0; k < 4 && X + k < W2; k++) d[k] = (uint8_t)v[k];
    }
}

cudaError_t launch_upsample2x(const uint8_t* d_src, size_t src_stride, int src_pitch, int w, int h, uint8_t* d_dst,
                              size_t dst_stride, int dst_pitch, int nframes, cudaStream_t st) {
    const int W2 = 2 * w - 2, H2 = 2 * h - 2;
    const dim3 grid((W2 / 4 + 256) / 256, H2, nframes), block(256);
    upsample2x_kernel<<<grid, block, 0, st>>>(d_src, src_stride, src_pitch, w, h, d_dst, dst_stride, dst_pitch);
    return cudaGetLastError();
}

cudaError_t launch_integral(const PipeP& P, const uint8_t* d_images, size_t image_stride, int pitch, int nframes,
                            int* d_integral, int* d_integral_ph, int* d_colsum, int* d_rowsum, int* d_tilesum, int* d_counts,
                            cudaStream_t st) {
    const dim3 grid(P.nchunks, P.nbands, nframes), block(kThreads);
    const bool aligned = (pitch % 8 == 0) && (image_stride % 8 == 0) && ((uintptr_t)d_images % 8 == 0);
    if (aligned) {
        integral_reduce<true><<<grid, block, 0, st>>>(P, d_images, image_stride, pitch, d_colsum, d_rowsum, d_tilesum, d_counts);
        return launch_dep(integral_scan<true>, grid, block, 0, st, P, d_images, image_stride, pitch, d_colsum, d_rowsum, d_tilesum, d_integral, d_integral_ph);
    } else {
        integral_reduce<false><<<grid, block, 0, st>>>(P, d_images, image_stride, pitch, d_colsum, d_rowsum, d_tilesum, d_counts);
        return launch_dep(integral_scan<false>, grid, block, 0, st, P, d_images, image_stride, pitch, d_colsum, d_rowsum, d_tilesum, d_integral, d_integral_ph);
    }
    return cudaGetLastError();
}

}  // namespace sb
