// tma_probe.cu -- what the TMA can do for the descriptor patches (run on the B200: tools/gpu_tma_probe.sh).
//   1. correctness of cp.async.bulk.tensor.3d boxes of the int32 integral image at UNALIGNED start columns, negative start
//      rows, row elementStride = step and the 64-byte / 128-byte swizzle (the layout the sweep reads with LDS.128);
//   2. throughput of such patch loads per SM (bytes landed in shared memory per second over the whole chip).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/tma_probe tools/tma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ int g_timeout;
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t it = 0; it < (1u << 16); it++) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
    }
    g_timeout = 1;
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

struct Cfg { int bw, K, step, nblk, nph, align, swz; };  // box width (ints), rows per box, row stride, column blocks, row phases

// one warp per CTA; loads a patch = nph phases x nblk column blocks, each box [K rows (stride step)][bw ints]
__global__ void probe_kernel(const __grid_constant__ CUtensorMap map, const int* __restrict__ img, int W, int H, int P, long long fstride,
                             Cfg c, int iters, int check, unsigned long long* bad, int* firstbad) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    const int lane = threadIdx.x;
    const uint32_t sbar = smem_u32(&bar);
    if (lane == 0) { mbar_init(sbar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    const int boxbytes = c.bw * 4 * c.K;
    const uint32_t sbase = smem_u32(smem);
    unsigned rng = blockIdx.x * 2654435761u + 12345u;
    unsigned long long nbad = 0;
    for (int it = 0; it < iters; it++) {
        rng = rng * 1664525u + 1013904223u;
        const int f = 0;
        int x0 = (int)((rng >> 8) % (unsigned)(W + 40)) - 20;   // unaligned, may start left of / run past the row
        rng = rng * 1664525u + 1013904223u;
        int y0 = (int)((rng >> 8) % (unsigned)(H + 40)) - 20;
        if (c.align == 1) x0 &= ~3;
        if (c.align == 2) { x0 = ((x0 < 0 ? 0 : x0) & ~15) % (W - 200); y0 = (y0 < 0 ? 0 : y0) % (H - 200); }
        if (g_timeout) break;
        if (lane == 0) {
            mbar_expect_tx(sbar, (uint32_t)(boxbytes * c.nblk * c.nph));
            for (int a = 0; a < c.nph; a++)
                for (int b = 0; b < c.nblk; b++)
                    tma_load_3d(sbase + (uint32_t)((a * c.nblk + b) * boxbytes), &map, sbar, x0 + b * c.bw, y0 + a, f);
        }
        mbar_wait(sbar, it & 1);
        if (check) {
            const int swz = !c.swz ? 0 : c.bw == 32 ? 7 : (c.bw == 16 ? 3 : 0);
            for (int a = 0; a < c.nph; a++)
                for (int b = 0; b < c.nblk; b++)
                    for (int k = 0; k < c.K; k++)
                        for (int x = lane; x < c.bw; x += 32) {
                            // row pitch bw*4 bytes; swizzle: 16-byte chunk index ^= (byte address >> 7) & swz
                            const int rowbyte = (a * c.nblk + b) * boxbytes + k * c.bw * 4;  // absolute: the pattern follows address bits
                            const int chunk = x >> 2;
                            const int pchunk = chunk ^ ((rowbyte >> 7) & swz);
                            const int got = *reinterpret_cast<const int*>(smem + rowbyte + pchunk * 16 + (x & 3) * 4);
                            const int gx = x0 + b * c.bw + x, gy = y0 + a + k * c.step;
                            const int want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? img[(size_t)f * fstride + (size_t)gy * P + gx] : 0;
                            if (got != want) { if (!nbad && atomicAdd(firstbad, 1) == 0) printf("bad: it %d a %d b %d k %d x %d got %d want %d (x0 %d y0 %d)\n", it, a, b, k, x, got, want, x0, y0); nbad++; }
                        }
            __syncwarp();
        }
    }
    if (nbad) atomicAdd(bad, nbad);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int W = 1921, H = 1083, P = 2048, F = 8;
    const long long fstride = (long long)P * H;
    std::vector<int> h((size_t)F * fstride);
    for (size_t i = 0; i < h.size(); i++) h[i] = (int)(i * 2654435761u >> 3);
    int* d; CK(cudaMalloc(&d, h.size() * 4)); CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    unsigned long long* bad; int* firstbad; CK(cudaMalloc(&bad, 8)); CK(cudaMalloc(&firstbad, 4));
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)p;
    if (argc < 8) { printf("usage: tma_probe bw K step nblk nph align swz [iters_check]\n"); return 2; }
    const Cfg cfgs[] = {{atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atoi(argv[7])}};
    for (const Cfg& c : cfgs) {
        CUtensorMap map;
        const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)F};
        const cuuint64_t strides[2] = {(cuuint64_t)P * 4, (cuuint64_t)fstride * 4};
        const cuuint32_t box[3] = {(cuuint32_t)c.bw, (cuuint32_t)(c.K * c.step), 1};
        const cuuint32_t estr[3] = {1, (cuuint32_t)c.step, 1};
        const CUtensorMapSwizzle sw = !c.swz ? CU_TENSOR_MAP_SWIZZLE_NONE : c.bw == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_INT32, 3, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("cfg bw %d K %d step %d nblk %d nph %d align %d swz %d: encode rc %d\n", c.bw, c.K, c.step, c.nblk, c.nph, c.align, c.swz, (int)r);
        if (r != CUDA_SUCCESS) continue;
        const int smem = c.bw * 4 * c.K * c.nblk * c.nph + 1024;
        CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaMemset(bad, 0, 8)); CK(cudaMemset(firstbad, 0, 4));
        probe_kernel<<<296, 32, smem>>>(map, d, W, H, P, fstride, c, 20, 1, bad, firstbad);
        CK(cudaDeviceSynchronize());
        unsigned long long hb; CK(cudaMemcpy(&hb, bad, 8, cudaMemcpyDeviceToHost));
        int to = 0; CK(cudaMemcpyFromSymbol(&to, g_timeout, 4));
        printf("  check: %llu mismatches, timeout %d\n", hb, to);
        if (to) { to = 0; CK(cudaMemcpyToSymbol(g_timeout, &to, 4)); continue; }
        for (int ctas_per_sm = 1; ctas_per_sm <= 8; ctas_per_sm *= 2) {
            int maxb = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxb, probe_kernel, 32, smem));
            if (ctas_per_sm > maxb) break;
            const int iters = 2000, grid = 148 * ctas_per_sm;
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            probe_kernel<<<grid, 32, smem>>>(map, d, W, H, P, fstride, c, 100, 0, bad, firstbad);
            CK(cudaEventRecord(e0));
            probe_kernel<<<grid, 32, smem>>>(map, d, W, H, P, fstride, c, iters, 0, bad, firstbad);
            CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            const double patches = (double)grid * iters, bytes = patches * (smem - 1024);
            printf("  %d warp(s)/SM, one patch in flight each: %.3f us per patch per warp, %.1f M patches/s, %.2f TB/s into smem (max resident %d)\n",
                   ctas_per_sm, 1e3 * ms / iters, patches / ms / 1e3, bytes / ms / 1e9, maxb);
        }
    }
    return 0;
}
