// match.cu -- brute-force descriptor matching on the 5th-generation tensor cores (tcgen05 + TMEM +
// TMA), with the reference's top-2 / ambiguity semantics fused into the epilogue.
//
// Replaces cuFindMaxCorr / findMaxCorr (surfd.cu:2535-2671, 3550-3566): S = F1 * F2^T, per row of
// F1 the best and "second" correlation and the five match fields of SurfPoint. It is the only
// dense contraction of the path. Reference semantics kept (bit-for-bit where they are observable):
//   * score = dot product accumulated as ONE fp32 FFMA chain over d = 0..nf-1 (surfd.cu:2591-2609);
//   * candidates are the first n2 - n2%32 descriptors of set 2 (surfd.cu:2569);
//   * candidates fall into 8 groups g = (p2 % 32) / 4, each keeping a running top-2 with strict >
//     from (0, 0, -1) in increasing p2 (surfd.cu:2610-2625); the merge starts from group 0 and sees
//     only the maxima of the other groups (surfd.cu:2646-2664); ambiguity = second / (best + 1e-6).
// Differences by design: writes are bounded to n1 rows (the reference writes whole 32-row blocks)
// and match_x/y stay 0 when nothing matched (the reference reads surf2[-1]).
//
// Three launches:
//   1. match_prep   fp32 descriptors -> bf16 "split" operands A' = [hi | hi | lo], B' = [hi | lo | hi]
//                   (K = 3*nf), so that A'.B'^T = hi.hi + hi.lo + lo.hi reproduces the fp32 dot product
//                   to ~2e-5 absolute on unit vectors (plain bf16 would be 4e-3: too coarse to rank).
//   2. match_mma    one CTA per (128-row block of A', range of column tiles of B'): TMA (128-byte
//                   swizzle) stages the operand tiles, ONE thread issues tcgen05.mma 128xBNx16 into a
//                   TMEM accumulator, the four warps read it back with tcgen05.ld and keep, per row and
//                   group, the running top-2 (value + index) -- the 10 M scores never touch memory.
//   3. match_final  per row: the 16 group candidates are re-scored with the reference's exact fp32
//                   FFMA chain, ordered with its update rule, merged with its merge rule and written.
//                   So score / match are those of the reference unless two candidates of one group
//                   are closer than the split-bf16 error to the group's second place.
#include <cuda.h>
#include <algorithm>
#include <stdlib.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace sb {

namespace {

constexpr int kBM = 128;        // rows of F1 per CTA == TMEM lanes
constexpr int kSwizzleRow = 128;  // bytes per operand row chunk (64 bf16) == one 128B-swizzle row

constexpr int kStages = 2;      // B' tiles in flight == TMEM accumulators
constexpr int kEpiWarps = 8;    // epilogue warps; warp kEpiWarps is the producer (TMA + MMA issue)

// MB = row blocks of A' per CTA: every B' tile that arrives in shared memory is multiplied with MB x 128 rows. The
// batched form uses MB = 2 for 64-d descriptors: with one row block a CTA streamed 48 KB of B' per 128 x 128 x 192 tile
// product, 1248 CTAs pulled 2.3 GB through L2 for 32 pairs and the kernel ran at the L2 -> SM bandwidth (ncu: 414 us,
// tensor pipe 38 % active).
template <int NF, int MB = 1> struct MatchCfg {
    static constexpr int KCH = 3 * NF / 64;            // 64-element K chunks of the split operands
    static constexpr int KTOT = 3 * NF;                // K of the tensor-core GEMM
    static constexpr int BN = NF == 64 ? 128 : 64;     // columns (descriptors of set 2) per tile: two stages must fit
    static constexpr int A_BYTES = KCH * kBM * kSwizzleRow;   // one row block
    static constexpr int B_BYTES = KCH * BN * kSwizzleRow;
    static constexpr int XCHG_BYTES = kBM * 8 * 16;  // per row block: top-2 of the upper column half, handed to the lower half's warps
    // (the exchange buffers reuse the B' stages: every tile has been multiplied when the first epilogue warp gets there)
    static_assert(MB * XCHG_BYTES <= kStages * B_BYTES, "exchange buffers alias the B' stages");
    static constexpr int SMEM = MB * A_BYTES + kStages * B_BYTES + 1024 /*alignment slack*/ + 128 /*barriers*/;
    static constexpr int TMEM_COLS = kStages * MB * BN;     // a power of two >= 32, <= 512
    static constexpr int THREADS = (kEpiWarps + 1) * 32;  // 8 epilogue warps (each scans its rows of all MB accumulators) + the producer
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a mis-programmed pipeline traps (clean launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 24); it++) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return;
    }
    __trap();
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
// Shared-memory matrix descriptor of a K-major operand tile stored as rows of 128 bytes with the
// 128-byte swizzle (what TMA writes): start>>4, LBO field 1, SBO = 1024 B between 8-row groups,
// descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor, kind::f16: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1), both K-major,
// N>>3 at bits 17-22, M>>4 at bits 24-28.
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

// Batched form (sb_match_pairs_async): pair z of the launch matches frame pairs[z].x against frame pairs[z].y of one
// detect batch, with the keypoint counts read ON THE DEVICE (no host round trip between detection and matching, and one
// launch sequence for all pairs instead of three launches per pair). counts == nullptr: the single-pair call, every
// per-pair quantity comes from the kernel arguments as before.
struct MatchBatch {
    const int* counts = nullptr;   // [frames]
    const int2* pairs = nullptr;   // [npairs]; nullptr: (2z, 2z + 1)
    sb_point* pts = nullptr;       // [frame][pts_stride]
    const float* desc = nullptr;   // [frame][desc_stride] floats
    long long pts_stride = 0, desc_stride = 0;
    int bound = 0;                 // counts are clamped to this; rows_cap = its multiple of 128 (rows per pair in the scratch)
    int rows_cap = 0;
    int nsplit = 1;
};
__device__ __forceinline__ int2 batch_pair(const MatchBatch& mb, int z) { return mb.pairs ? mb.pairs[z] : make_int2(2 * z, 2 * z + 1); }
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

// ---------------------------------------------------------------------------- 1. operand split
// rows >= nvalid are zero (padding of A', non-candidates of B'); out: [rows_pad][3*NF] bf16. One launch for both sets.
template <int NF>
__global__ void match_prep(const float* __restrict__ fa, int na, int na_pad, __nv_bfloat16* __restrict__ outa,
                           const float* __restrict__ fb, int nb, int nb_pad, __nv_bfloat16* __restrict__ outb,
                           const __grid_constant__ MatchBatch mb) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // match_mma's prologue overlaps this kernel
    // a thread converts 8 consecutive elements of a row: two 128-bit loads, three 128-bit stores (hi, and lo / hi again)
    constexpr int PER = NF / 8;
    constexpr int BN = NF == 64 ? 128 : 64;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    int row = t / PER;
    const int d = (t - row * PER) * 8;
    bool is_b = row >= na_pad;
    if (mb.counts) {
        // blockIdx.y = 2 * pair + role; rows [0, rows_cap) of the pair's slot: valid rows, then zeros up to the padded size
        const int z = blockIdx.y >> 1;
        is_b = blockIdx.y & 1;
        const int2 pr = batch_pair(mb, z);
        const int fr = is_b ? pr.y : pr.x;
        const int n = min(mb.counts[fr], mb.bound);
        if (is_b) { nb = n - (n & 31); nb_pad = max((nb + BN - 1) / BN, 1) * BN; }
        else { na = n; na_pad = (n + kBM - 1) / kBM * kBM; }
        if (row >= (is_b ? nb_pad : na_pad)) return;
        fa = fb = mb.desc + (size_t)fr * mb.desc_stride;
        outa += (size_t)z * mb.rows_cap * (3 * NF);
        outb += (size_t)z * mb.rows_cap * (3 * NF);
    } else {
        if (is_b) row -= na_pad;
        if (is_b && row >= nb_pad) return;
    }
    const float* f = is_b ? fb : fa;
    const int nvalid = is_b ? nb : na;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (row < nvalid) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(f + (size_t)row * NF + d));
        const float4 y = __ldg(reinterpret_cast<const float4*>(f + (size_t)row * NF + d) + 1);
        v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
    }
    __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        hi[k] = __float2bfloat16_rn(v[k]);
        lo[k] = __float2bfloat16_rn(v[k] - __bfloat162float(hi[k]));
    }
    __nv_bfloat16* o = (is_b ? outb : outa) + (size_t)row * (3 * NF) + d;
    const uint4 h4 = *reinterpret_cast<const uint4*>(hi), l4 = *reinterpret_cast<const uint4*>(lo);
    *reinterpret_cast<uint4*>(o) = h4;
    *reinterpret_cast<uint4*>(o + NF) = is_b ? l4 : h4;
    *reinterpret_cast<uint4*>(o + 2 * NF) = is_b ? h4 : l4;
}

// ---------------------------------------------------------------------------- 2. tensor-core scores + fused group top-2
struct __align__(16) Top2 { float mx, sc; int imx, isc; };  // 16 bytes, moved as one 128-bit word

// Warp-specialised and double-buffered: warp 8 (one thread) runs the operand pipeline -- TMA of tile i+1 into the
// other shared-memory stage, tcgen05.mma of tile i into the other TMEM accumulator -- while warps 0-7 run the top-2
// epilogue of tile i-1 out of TMEM. Per stage three mbarriers: `full` (TMA landed), `mma` (tcgen05.commit: accumulator
// ready, operands consumed), `free` (the eight epilogue warps have read the accumulator). Round-1 v1 did load -> MMA ->
// epilogue strictly one after the other (ncu: tensor pipe active 12.5 % of the kernel).
template <int NF, bool KEYS, int MB>
__global__ void __launch_bounds__((kEpiWarps + 1) * 32, 1)
match_mma(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int ntiles, int tiles_per_split,
          int n1pad, Top2* __restrict__ part, const __grid_constant__ MatchBatch mb) {
    using Cfg = MatchCfg<NF, MB>;
    constexpr int BN = Cfg::BN, KCH = Cfg::KCH;
    const int zpair = blockIdx.z;
    if (mb.counts) {
        // per-pair geometry from the device-side counts; row blocks past the pair's rows leave at once
        const int2 pr = batch_pair(mb, zpair);
        const int n1 = min(mb.counts[pr.x], mb.bound), n2 = min(mb.counts[pr.y], mb.bound);
        if ((int)blockIdx.x * MB * kBM >= n1) return;
        const int ncand = n2 - (n2 & 31);
        ntiles = (ncand + BN - 1) / BN;
        tiles_per_split = (ntiles + mb.nsplit - 1) / mb.nsplit;
        n1pad = mb.rows_cap;
        part += (size_t)zpair * mb.nsplit * mb.rows_cap * 8;
    }
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                         // MB x (KCH chunks of 128 rows x 128 B)
    uint8_t* sB = smem + MB * Cfg::A_BYTES;     // kStages x (KCH chunks of BN rows x 128 B)
    float4* xchg = reinterpret_cast<float4*>(sB);  // after the last tile
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MB * Cfg::A_BYTES + kStages * Cfg::B_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kStages);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // match_final's CTAs may take their places now
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rb0 = blockIdx.x * MB, split = blockIdx.y;
    const int t0 = split * tiles_per_split, t1 = min(ntiles, t0 + tiles_per_split);
    const int n = max(t1 - t0, 0);
    auto bar_full = [&](int st) { return smem_u32(&bars[st]); };
    auto bar_mma = [&](int st) { return smem_u32(&bars[kStages + st]); };
    auto bar_free = [&](int st) { return smem_u32(&bars[2 * kStages + st]); };

    constexpr uint32_t idesc = umma_idesc(kBM, BN);
    auto issue_tma = [&](int i) {  // tile t0+i into stage i % kStages (and, once, the CTA's rows of A')
        const int st = i % kStages;
        mbar_expect_tx(bar_full(st), Cfg::B_BYTES + (i == 0 ? MB * Cfg::A_BYTES : 0));
        // (the maps are [pair][row][k]: the single-pair call has one pair; rows past the operand buffer arrive as zeros)
        if (i == 0)
            for (int m = 0; m < MB; m++)
                for (int c = 0; c < KCH; c++)
                    tma_load_3d(smem_u32(sA + m * Cfg::A_BYTES + c * kBM * kSwizzleRow), &mapA, bar_full(st), c * 64, (rb0 + m) * kBM, zpair);
        for (int c = 0; c < KCH; c++)
            tma_load_3d(smem_u32(sB + st * Cfg::B_BYTES + c * BN * kSwizzleRow), &mapB, bar_full(st), c * 64, (t0 + i) * BN, zpair);
    };
    if (tid == kEpiWarps * 32) {
        // the producer thread arms the barriers itself and starts the first loads at once: they fly while warp 1 allocates
        // tensor memory and the CTA meets at the barrier below
        for (int st = 0; st < kStages; st++) { mbar_init(bar_full(st), 1); mbar_init(bar_mma(st), 1); mbar_init(bar_free(st), kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        // (programmatic dependent launch: everything above overlaps match_prep; its output is first touched here)
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if (n > 0) issue_tma(0);
    }
    if (warp == 1) {  // TMEM: kStages x MB accumulators of BN fp32 columns x 128 lanes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;

    if (warp == kEpiWarps) {
        // ------------------------------------------------------------------ producer: one thread
        if (lane == 0 && n > 0) {
            for (int i = 0; i < n; i++) {
                const int st = i % kStages, use = i / kStages;
                if (i + 1 < n) {
                    // stage (i+1) % kStages was last read by the MMAs of tile i+1-kStages
                    if (i + 1 >= kStages) mbar_wait(bar_mma((i + 1) % kStages), ((i + 1) / kStages - 1) & 1);
                    issue_tma(i + 1);
                }
                mbar_wait(bar_full(st), use & 1);
                if (use >= 1) mbar_wait(bar_free(st), (use - 1) & 1);  // the epilogue of tile i-kStages has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // D_m[128 x BN] = sum over K chunks and 16-element K steps, for each of the CTA's row blocks
                for (int m = 0; m < MB; m++)
                    for (int c = 0; c < KCH; c++) {
                        const uint64_t da = umma_desc(smem_u32(sA + m * Cfg::A_BYTES + c * kBM * kSwizzleRow));
                        const uint64_t db = umma_desc(smem_u32(sB + st * Cfg::B_BYTES + c * BN * kSwizzleRow));
#pragma unroll
                        for (int k = 0; k < 4; k++)  // +32 bytes (2 x 16 B units) per K step inside the swizzle row
                            umma_bf16(tmem_d + (st * MB + m) * BN, da + 2 * k, db + 2 * k, idesc, (c | k) ? 1u : 0u);
                    }
                umma_commit(bar_mma(st));  // arrives when the MMAs above have completed (implies before_thread_sync)
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue: 8 warps
        const int quarter = warp & 3;            // TMEM lane quarter this warp may read (warp id % 4)
        const int chalf = (warp >> 2) & 1;       // which half of the tile's columns this warp scans
        // per row block m of the CTA: the thread's row is TMEM lane quarter*32 + lane of accumulator (stage, m)
        float mx[MB][8], sc[MB][8];
        int imx[MB][8], isc[MB][8];
#pragma unroll
        for (int m = 0; m < MB; m++)
#pragma unroll
            for (int g = 0; g < 8; g++) { mx[m][g] = 0.f; sc[m][g] = 0.f; imx[m][g] = -1; isc[m][g] = -1; }
        // Inside a round of up to kRound tiles the running top-2 of a group is kept on PACKED KEYS: the score's bit pattern
        // with its low 6 mantissa bits replaced by 63 - (candidate number inside the round). Positive floats order like
        // signed integers, so a score costs one LOP3 and three integer min/max instead of two compares, three selects
        // and three float min/max (ncu of round 1: the epilogue's 9 instructions per score, not the tensor pipe, set the
        // kernel's time). The key drops 2^-17 of the score -- below the 2e-5 of the split-bf16 product, and like it only
        // a matter of WHICH two candidates match_final re-scores exactly; equal truncated scores prefer the lower index,
        // as the reference's strict > does. Key 0 = (score 0, candidate 63) is the empty slot.
        constexpr int kRound = 63 / (BN / 16);  // BN/16 candidates of one group per tile and thread (BN/2 columns / 8 groups); code 63 = empty
        int k1[MB][8], k2[MB][8];
        auto fold = [&](int round0) {  // keys of the round starting at tile round0 -> (value, index), merged into the running top-2
#pragma unroll
            for (int m = 0; m < MB; m++)
#pragma unroll
                for (int g = 0; g < 8; g++) {
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        const int key = q == 0 ? k1[m][g] : k2[m][g];
                        const int code = 63 - (key & 63);
                        if (code != 63) {
                            // code = tile-in-round * (BN/16) + chunk * 4 + column-in-group
                            const int per_tile = BN / 16;
                            const int ti = code / per_tile, rest = code - ti * per_tile;
                            const int col = (t0 + round0 + ti) * BN + chalf * (BN / 2) + (rest >> 2) * 32 + g * 4 + (rest & 3);
                            const float s = __int_as_float(key & ~63);
                            if (s > mx[m][g]) { sc[m][g] = mx[m][g]; isc[m][g] = imx[m][g]; mx[m][g] = s; imx[m][g] = col; }
                            else if (s > sc[m][g]) { sc[m][g] = s; isc[m][g] = col; }
                        }
                    }
                    k1[m][g] = 0; k2[m][g] = 0;
                }
        };
#pragma unroll
        for (int m = 0; m < MB; m++)
#pragma unroll
            for (int g = 0; g < 8; g++) { k1[m][g] = 0; k2[m][g] = 0; }
        int round0 = 0;
        for (int i = 0; i < n; i++) {
            const int st = i % kStages, use = i / kStages;
            if (KEYS && i - round0 == kRound) { fold(round0); round0 = i; }
            mbar_wait(bar_mma(st), use & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // thread == (row = TMEM lane, half of the columns). All of this warp's columns of one accumulator are read
            // first (one wait), then scanned; after the read of the LAST accumulator of the stage it goes back to the MMA
            // issuer. The 8 groups of a 32-column chunk are independent dependency chains (ILP 8).
            constexpr int NCH = BN / 2 / 32;  // 32-column chunks per warp and accumulator: 2 (BN = 128) or 1 (BN = 64)
#pragma unroll
            for (int m = 0; m < MB; m++) {
                uint32_t r[NCH][32];
#pragma unroll
                for (int c = 0; c < NCH; c++)
                    tmem_ld32(tmem_d + ((uint32_t)(quarter * 32) << 16) + (st * MB + m) * BN + chalf * (BN / 2) + 32 * c, r[c]);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (m == MB - 1) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_free(st)) : "memory");
                }
#pragma unroll
                for (int c = 0; c < NCH; c++) {
                    // 63 - candidate number of column-in-group 0 of this chunk; the three others follow downwards
                    const int code0 = 63 - ((i - round0) * (BN / 16) + c * 4);
                    if (KEYS) {
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            const int g = j >> 2;
                            const int key = (int)(r[c][j] & 0xFFFFFFC0u) | (code0 - (j & 3));
                            const int t = min(k1[m][g], key);
                            k1[m][g] = max(k1[m][g], key);
                            k2[m][g] = max(k2[m][g], t);
                        }
                    } else {
                        const int col0 = (t0 + i) * BN + chalf * (BN / 2) + 32 * c;  // multiple of 32: group of column j is (j % 32) / 4
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            const int g = j >> 2;
                            const float s = __uint_as_float(r[c][j]);
                            const bool gt1 = s > mx[m][g], gt2 = s > sc[m][g];
                            isc[m][g] = gt1 ? imx[m][g] : (gt2 ? col0 + j : isc[m][g]);
                            imx[m][g] = gt1 ? col0 + j : imx[m][g];
                            sc[m][g] = fmaxf(sc[m][g], fminf(mx[m][g], s));
                            mx[m][g] = fmaxf(mx[m][g], s);
                        }
                    }
                }
            }
        }
        if (KEYS) fold(round0);
        // The two column halves of a row meet here: the warps of the upper half pass their top-2 through shared memory
        // (a named barrier over the 8 epilogue warps), the lower half's warps merge -- strict >, so equal scores keep
        // the lower index, which is theirs -- and write ONE entry per (split, row, group).
        const int row = quarter * 32 + lane;
        if (chalf == 1) {
#pragma unroll
            for (int m = 0; m < MB; m++)
#pragma unroll
                for (int g = 0; g < 8; g++)
                    xchg[(m * kBM + row) * 8 + g] = make_float4(mx[m][g], sc[m][g], __int_as_float(imx[m][g]), __int_as_float(isc[m][g]));
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
        if (chalf == 0) {
#pragma unroll
            for (int m = 0; m < MB; m++) {
                if ((rb0 + m) * kBM >= n1pad) continue;  // (a CTA's last row block may lie past the rows)
                Top2* dst = part + ((size_t)split * n1pad + (size_t)(rb0 + m) * kBM + row) * 8;
#pragma unroll
                for (int g = 0; g < 8; g++) {
                    const float4 o = xchg[(m * kBM + row) * 8 + g];
                    const float ov[2] = {o.x, o.y};
                    const int oi[2] = {__float_as_int(o.z), __float_as_int(o.w)};
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        if (oi[q] < 0) continue;
                        if (ov[q] > mx[m][g]) { sc[m][g] = mx[m][g]; isc[m][g] = imx[m][g]; mx[m][g] = ov[q]; imx[m][g] = oi[q]; }
                        else if (ov[q] > sc[m][g]) { sc[m][g] = ov[q]; isc[m][g] = oi[q]; }
                    }
                    reinterpret_cast<float4*>(dst)[g] = make_float4(mx[m][g], sc[m][g], __int_as_float(imx[m][g]), __int_as_float(isc[m][g]));
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(Cfg::TMEM_COLS) : "memory");
}

// ---------------------------------------------------------------------------- 3. exact re-score, group rule, merge
template <int NF>
__device__ __forceinline__ float exact_dot(const float* __restrict__ a, const float* __restrict__ b) {
    float s = 0.f;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
#pragma unroll 8
    for (int d = 0; d < NF / 4; d++) {
        const float4 x = __ldg(a4 + d), y = __ldg(b4 + d);
        s = __fmaf_rn(x.x, y.x, s); s = __fmaf_rn(x.y, y.y, s); s = __fmaf_rn(x.z, y.z, s); s = __fmaf_rn(x.w, y.w, s);
    }
    return s;
}

// One WARP per row of set 1. Lane (2g + k), k in {0,1}, owns candidate k of group g: it picks the
// k-th best tensor-core score of the group across the column splits, re-scores it with the
// reference's exact fp32 FFMA chain (a chain cannot be split across lanes), and the pair is
// combined with the reference's running update in index order; lane 0 then applies the merge.
template <int NF>
__global__ void __launch_bounds__(128)
match_final(sb_point* __restrict__ pts1, int n1, const float* __restrict__ f1, const sb_point* __restrict__ pts2,
            const float* __restrict__ f2, const Top2* __restrict__ part, int nsplit, int n1pad, const __grid_constant__ MatchBatch mb) {
    // TWO rows per warp, one per half-warp (the 16 group candidates of a row are one lane each)
    const int lane = threadIdx.x & 31, hbase = lane & 16;
    const int p1 = 2 * (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) + (lane >> 4);
    asm volatile("griddepcontrol.wait;" ::: "memory");  // launched early (programmatic dependent launch): match_mma's partials
    if (mb.counts) {
        const int z = blockIdx.y;
        const int2 pr = batch_pair(mb, z);
        n1 = min(mb.counts[pr.x], mb.bound);
        pts1 = mb.pts + (size_t)pr.x * mb.pts_stride; pts2 = mb.pts + (size_t)pr.y * mb.pts_stride;
        f1 = mb.desc + (size_t)pr.x * mb.desc_stride; f2 = mb.desc + (size_t)pr.y * mb.desc_stride;
        nsplit = mb.nsplit; n1pad = mb.rows_cap;
        part += (size_t)z * mb.nsplit * mb.rows_cap * 8;
    }
    if ((p1 & ~1) >= n1) return;  // both rows of the warp
    const bool valid = p1 < n1;
    const int g = (lane >> 1) & 7, k = lane & 1;
    // tensor-core top-2 of the group across the splits. The splits cover increasing column ranges and are taken in
    // order, so strict > keeps the lower index on equal scores, as a running scan would.
    float v1 = 0.f, v2 = 0.f;
    int i1 = -1, i2 = -1;
    // (loads of eight splits are issued together: one L2 round trip per eight, not per split)
    for (int s0 = 0; s0 < nsplit; s0 += 8) {
        float4 cs[8];
#pragma unroll
        for (int u = 0; u < 8; u++)  // 128 contiguous bytes per (split,row) across the warp, one 128-bit load per lane
            cs[u] = s0 + u < nsplit ? __ldg(reinterpret_cast<const float4*>(part + ((size_t)(s0 + u) * n1pad + p1) * 8 + g))
                                    : make_float4(0.f, 0.f, __int_as_float(-1), __int_as_float(-1));
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const float cv[2] = {cs[u].x, cs[u].y};
            const int ci[2] = {__float_as_int(cs[u].z), __float_as_int(cs[u].w)};
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const bool ok = ci[q] >= 0;
                const bool gt1 = ok && cv[q] > v1, gt2 = ok && cv[q] > v2;
                i2 = gt1 ? i1 : (gt2 ? ci[q] : i2);
                v2 = gt1 ? v1 : (gt2 ? cv[q] : v2);
                i1 = gt1 ? ci[q] : i1;
                v1 = gt1 ? cv[q] : v1;
            }
        }
    }
    // candidate of this lane: the pair in increasing index order (k = 0 first)
    int lo = i1, hi = i2;
    if (lo < 0 || (hi >= 0 && hi < lo)) { const int t = lo; lo = hi; hi = t; }
    const int mine = k == 0 ? lo : hi;
    // Only candidates that can reach the result are re-scored exactly. The merge below returns m = the largest group
    // maximum and s2 = the larger of group 0's second and the second largest group maximum, whatever the order of the
    // groups; a candidate whose tensor-core score lies more than kMargin (30 x the 3e-5 ranking resolution) below the
    // second largest group maximum can be neither, so it enters the merge with its approximate score. That leaves 2-3
    // of the 16 candidates of a row: 16 exact re-scores per row were 4 KB of descriptor reads from L2 per row, which
    // bound this kernel in the batched form (ncu: 197 us for 32 pairs of 4.9 k rows).
    constexpr float kMargin = 1e-3f;
    float t1 = v1, t2 = -1.f;  // top-2 of the groups' (approximate) maxima over the 8 groups of this half-warp
#pragma unroll
    for (int o = 2; o < 16; o <<= 1) {
        const float a = __shfl_xor_sync(0xffffffffu, t1, o), b = __shfl_xor_sync(0xffffffffu, t2, o);
        t2 = fmaxf(fminf(t1, a), fmaxf(t2, b));
        t1 = fmaxf(t1, a);
    }
    const float approx = mine == i1 ? v1 : v2;
    float e = approx;
    if (valid && mine >= 0 && approx >= t2 - kMargin) e = exact_dot<NF>(f1 + (size_t)p1 * NF, f2 + (size_t)mine * NF);
    // the reference's running update (surfd.cu:2610-2625) over (first, second) in index order
    const float e0 = __shfl_sync(0xffffffffu, e, lane & ~1), e1 = __shfl_sync(0xffffffffu, e, lane | 1);
    const int c0 = __shfl_sync(0xffffffffu, mine, lane & ~1), c1 = __shfl_sync(0xffffffffu, mine, lane | 1);
    float gm = 0.f, gs = 0.f;
    int gi = -1;
    if (c0 >= 0 && e0 > gm) { gm = e0; gi = c0; }
    if (c1 >= 0) {
        if (e1 > gm) { gs = gm; gm = e1; gi = c1; }
        else if (e1 > gs) gs = e1;
    }
    // merge (surfd.cu:2646-2664): start from group 0, the other groups contribute their maxima only
    float m = __shfl_sync(0xffffffffu, gm, hbase), s2 = __shfl_sync(0xffffffffu, gs, hbase);
    int idx = __shfl_sync(0xffffffffu, gi, hbase);
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const float qm = __shfl_sync(0xffffffffu, gm, hbase + 2 * q);
        const int qi = __shfl_sync(0xffffffffu, gi, hbase + 2 * q);
        if (idx != qi) {
            if (qm > m) { s2 = fmaxf(m, s2); m = qm; idx = qi; }
            else if (qm > s2) s2 = qm;
        }
    }
    if (valid && (lane & 15) == 0) {
        sb_point* p = pts1 + p1;
        p->score = m;
        p->match = idx;
        p->match_x = idx >= 0 ? pts2[idx].x : 0.f;
        p->match_y = idx >= 0 ? pts2[idx].y : 0.f;
        p->ambiguity = __fdiv_rn(s2, __fadd_rn(m, 1e-6f));
    }
}

// ---------------------------------------------------------------------------- descriptor sizes other than 64 / 128
//
// desc_wsz < 4 gives 16-, 32-, 36- or 72-d descriptors (surf.cpp:78-79), which the 64-element K chunks of the tensor-core
// path do not tile. Those sets are small problems (K <= 72); a warp per row of set 1 computes the reference's exact fp32
// FFMA chain against every candidate: lane l takes p2 = l, l+32, ... -- group (p2 % 32) / 4 = l / 4 -- so a lane's own
// running top-2 is a sub-sequence of its group's scan, four lanes merge into the group's (max, first arg max, second),
// and the eight groups merge with the reference's rule (surfd.cu:2646-2664).
__global__ void __launch_bounds__(128)
match_generic_kernel(sb_point* __restrict__ pts1, int n1, const float* __restrict__ f1, const sb_point* __restrict__ pts2, int ncand,
                     const float* __restrict__ f2, int nf) {
    extern __shared__ float s_row[];  // [4 warps][nf]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p1 = blockIdx.x * 4 + warp;
    if (p1 >= n1) return;
    float* a = s_row + warp * nf;
    for (int d = lane; d < nf; d += 32) a[d] = f1[(size_t)p1 * nf + d];
    __syncwarp();
    float mx = 0.f, sc = 0.f;
    int id = -1;
    for (int p2 = lane; p2 < ncand; p2 += 32) {
        const float* b = f2 + (size_t)p2 * nf;
        float s = 0.f;
        for (int d = 0; d < nf; d++) s = __fmaf_rn(a[d], __ldg(b + d), s);
        if (s > mx) { sc = mx; mx = s; id = p2; }
        else if (s > sc) sc = s;
    }
    // group = 4 consecutive lanes: the scan in increasing p2 keeps the first arg max and the second largest value
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
        const float omx = __shfl_xor_sync(0xffffffffu, mx, o), osc = __shfl_xor_sync(0xffffffffu, sc, o);
        const int oid = __shfl_xor_sync(0xffffffffu, id, o);
        if (oid >= 0 && (omx > mx || (omx == mx && (id < 0 || oid < id)))) { sc = fmaxf(fmaxf(mx, sc), osc); mx = omx; id = oid; }
        else sc = fmaxf(sc, fmaxf(oid >= 0 ? omx : 0.f, osc));
    }
    // merge (surfd.cu:2646-2664): start from group 0, the other groups contribute their maxima only
    float m = __shfl_sync(0xffffffffu, mx, 0), s2 = __shfl_sync(0xffffffffu, sc, 0);
    int idx = __shfl_sync(0xffffffffu, id, 0);
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const float qm = __shfl_sync(0xffffffffu, mx, 4 * q);
        const int qi = __shfl_sync(0xffffffffu, id, 4 * q);
        if (idx != qi) {
            if (qm > m) { s2 = fmaxf(m, s2); m = qm; idx = qi; }
            else if (qm > s2) s2 = qm;
        }
    }
    if (lane == 0) {
        sb_point* p = pts1 + p1;
        p->score = m;
        p->match = idx;
        p->match_x = idx >= 0 ? pts2[idx].x : 0.f;
        p->match_y = idx >= 0 ? pts2[idx].y : 0.f;
        p->ambiguity = __fdiv_rn(s2, __fadd_rn(m, 1e-6f));
    }
}

// ---------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// [rows][ktot] bf16 row-major; box = 64 elements (128 B) x box_rows, 128-byte swizzle
bool make_map(CUtensorMap* m, const void* base, int rows, int ktot, int box_rows, int npairs = 1) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)ktot, (cuuint64_t)rows, (cuuint64_t)npairs};
    const cuuint64_t strides[2] = {(cuuint64_t)ktot * 2, (cuuint64_t)ktot * 2 * (cuuint64_t)rows};
    const cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int NF>
cudaError_t run_match(sb_point* d_pts1, int n1, const float* d_f1, const sb_point* d_pts2, int n2, const float* d_f2,
                      MatchScratch& ws, int sm_count, cudaStream_t st) {
    using Cfg = MatchCfg<NF, 1>;
    const int ncand = n2 - (n2 & 31);
    const int n1pad = (n1 + kBM - 1) / kBM * kBM;
    const int ntiles = (ncand + Cfg::BN - 1) / Cfg::BN;
    const int n2pad = max(ntiles, 1) * Cfg::BN;
    const int rbs = n1pad / kBM;
    // about one CTA per SM (one wave): a CTA's fixed cost (TMEM allocation, its A' rows) is amortised over its tiles
    int nsplit = ntiles > 0 ? min(ntiles, max(1, sm_count / rbs)) : 1;
    const int tiles_per_split = ntiles > 0 ? (ntiles + nsplit - 1) / nsplit : 0;
    if (ntiles > 0) nsplit = (ntiles + tiles_per_split - 1) / tiles_per_split;
    // scratch (grow-only)
    const size_t needA = (size_t)n1pad * Cfg::KTOT * 2, needB = (size_t)n2pad * Cfg::KTOT * 2;
    const size_t needP = (size_t)nsplit * n1pad * 8 * sizeof(Top2);
    cudaError_t e;
    auto grow = [&](void*& p, size_t& cap, size_t need) -> cudaError_t {
        if (cap >= need) return cudaSuccess;
        if (p) { cudaError_t r = cudaFree(p); if (r != cudaSuccess) return r; p = nullptr; cap = 0; }
        cudaError_t r = cudaMalloc(&p, need);
        if (r == cudaSuccess) cap = need;
        return r;
    };
    if ((e = grow(ws.a, ws.cap_a, needA)) != cudaSuccess) return e;
    if ((e = grow(ws.b, ws.cap_b, needB)) != cudaSuccess) return e;
    if ((e = grow(ws.part, ws.cap_part, needP)) != cudaSuccess) return e;

    const MatchBatch none;
    match_prep<NF><<<((n1pad + n2pad) * (NF / 8) + 255) / 256, 256, 0, st>>>(d_f1, n1, n1pad, (__nv_bfloat16*)ws.a, d_f2, ncand, n2pad, (__nv_bfloat16*)ws.b, none);
    static_assert(sizeof(CUtensorMap) == sizeof(ws.map_a), "tensor map size");
    CUtensorMap& mapA = *reinterpret_cast<CUtensorMap*>(ws.map_a);
    CUtensorMap& mapB = *reinterpret_cast<CUtensorMap*>(ws.map_b);
    if (ws.map_nf != NF || ws.map_pairs != 1 || ws.map_a_base != ws.a || ws.map_a_rows != n1pad) {
        if (!make_map(&mapA, ws.a, n1pad, Cfg::KTOT, kBM)) return cudaErrorNotSupported;
        ws.map_a_base = ws.a; ws.map_a_rows = n1pad;
    }
    if (ws.map_nf != NF || ws.map_pairs != 1 || ws.map_b_base != ws.b || ws.map_b_rows != n2pad) {
        if (!make_map(&mapB, ws.b, n2pad, Cfg::KTOT, Cfg::BN)) return cudaErrorNotSupported;
        ws.map_b_base = ws.b; ws.map_b_rows = n2pad;
    }
    ws.map_nf = NF; ws.map_pairs = 1;
    // per device (a process may hold contexts on several GPUs), so set on every launch
    static const int use_pdl = getenv("SB_MATCH_PDL") ? atoi(getenv("SB_MATCH_PDL")) : 2;  // 0 off, 1 both, 2 mma only
    static const bool use_keys = !getenv("SB_MATCH_KEYS") || atoi(getenv("SB_MATCH_KEYS")) != 0;
    if ((e = cudaFuncSetAttribute(match_mma<NF, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(match_mma<NF, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM)) != cudaSuccess) return e;
    // match_mma is a programmatic dependent of match_prep: its CTAs are placed and run their prologue (barrier set-up,
    // tensor-memory allocation) while match_prep is still running, and the producer thread waits at griddepcontrol.wait
    // before the first TMA touches match_prep's output. match_final is launched normally: as a programmatic dependent its
    // 685 CTAs sat on the SMs beside match_mma's and the three kernels took 24.7 us instead of 20.5 (B200, CUDA graph of
    // the three launches, 2739 x 3443; no dependent launch at all: 20.8). SB_MATCH_PDL / SB_MATCH_KEYS are measurement
    // switches.
    cudaLaunchAttribute pdl[1];
    pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pdl[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(rbs, nsplit); cfg.blockDim = dim3((kEpiWarps + 1) * 32); cfg.dynamicSmemBytes = Cfg::SMEM; cfg.stream = st;
    cfg.attrs = pdl; cfg.numAttrs = use_pdl ? 1 : 0;
    if (use_keys) e = cudaLaunchKernelEx(&cfg, match_mma<NF, true, 1>, mapA, mapB, ntiles, tiles_per_split, n1pad, (Top2*)ws.part, none);
    else e = cudaLaunchKernelEx(&cfg, match_mma<NF, false, 1>, mapA, mapB, ntiles, tiles_per_split, n1pad, (Top2*)ws.part, none);
    if (e != cudaSuccess) return e;
    cfg.gridDim = dim3((n1 + 7) / 8); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 0;
    cfg.numAttrs = use_pdl == 1 ? 1 : 0;
    if ((e = cudaLaunchKernelEx(&cfg, match_final<NF>, d_pts1, n1, d_f1, d_pts2, d_f2, (const Top2*)ws.part, nsplit, n1pad, none)) != cudaSuccess) return e;
    return cudaGetLastError();
}

// All pairs of a detect batch in one launch sequence; counts stay on the device. Scratch: rows_cap rows per pair.
template <int NF>
cudaError_t run_match_batch(sb_point* d_pts, long long pts_stride, const int* d_counts, const float* d_desc, long long desc_stride,
                            int npairs, const int2* d_pairs, int bound, MatchScratch& ws, int sm_count, cudaStream_t st) {
    constexpr int MB = NF == 64 ? 2 : 1;  // two row blocks per CTA where the shared memory allows
    using Cfg = MatchCfg<NF, MB>;
    MatchBatch mb;
    mb.counts = d_counts; mb.pairs = d_pairs; mb.pts = d_pts; mb.desc = d_desc; mb.pts_stride = pts_stride; mb.desc_stride = desc_stride;
    mb.bound = bound;
    mb.rows_cap = (bound + kBM - 1) / kBM * kBM;   // a multiple of 128, hence of BN
    const int rbs_cap = (mb.rows_cap / kBM + MB - 1) / MB;  // CTAs along the rows
    // enough CTAs for about two waves; with many pairs every CTA keeps its A' rows for ALL column tiles (no split)
    mb.nsplit = std::max(1, std::min(8, (2 * sm_count + npairs * rbs_cap - 1) / (npairs * rbs_cap)));
    const size_t needO = (size_t)npairs * mb.rows_cap * Cfg::KTOT * 2;
    const size_t needP = (size_t)npairs * mb.nsplit * mb.rows_cap * 8 * sizeof(Top2);
    cudaError_t e;
    auto grow = [&](void*& p, size_t& cap, size_t need) -> cudaError_t {
        if (cap >= need) return cudaSuccess;
        if (p) { cudaError_t r = cudaFree(p); if (r != cudaSuccess) return r; p = nullptr; cap = 0; }
        cudaError_t r = cudaMalloc(&p, need);
        if (r == cudaSuccess) cap = need;
        return r;
    };
    if ((e = grow(ws.a, ws.cap_a, needO)) != cudaSuccess) return e;
    if ((e = grow(ws.b, ws.cap_b, needO)) != cudaSuccess) return e;
    if ((e = grow(ws.part, ws.cap_part, needP)) != cudaSuccess) return e;
    CUtensorMap& mapA = *reinterpret_cast<CUtensorMap*>(ws.map_a);
    CUtensorMap& mapB = *reinterpret_cast<CUtensorMap*>(ws.map_b);
    if (ws.map_nf != NF || ws.map_pairs != npairs || ws.map_a_base != ws.a || ws.map_a_rows != mb.rows_cap ||
        ws.map_b_base != ws.b || ws.map_b_rows != mb.rows_cap) {
        if (!make_map(&mapA, ws.a, mb.rows_cap, Cfg::KTOT, kBM, npairs) || !make_map(&mapB, ws.b, mb.rows_cap, Cfg::KTOT, Cfg::BN, npairs))
            return cudaErrorNotSupported;
        ws.map_a_base = ws.a; ws.map_b_base = ws.b; ws.map_a_rows = ws.map_b_rows = mb.rows_cap; ws.map_nf = NF; ws.map_pairs = npairs;
    }
    match_prep<NF><<<dim3((mb.rows_cap * (NF / 8) + 255) / 256, 2 * npairs), 256, 0, st>>>(nullptr, 0, 0, (__nv_bfloat16*)ws.a, nullptr, 0, 0,
                                                                                          (__nv_bfloat16*)ws.b, mb);
    if ((e = cudaFuncSetAttribute(match_mma<NF, true, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM)) != cudaSuccess) return e;
    match_mma<NF, true, MB><<<dim3(rbs_cap, mb.nsplit, npairs), Cfg::THREADS, Cfg::SMEM, st>>>(mapA, mapB, 0, 0, 0, (Top2*)ws.part, mb);
    match_final<NF><<<dim3(mb.rows_cap / 8, npairs), 128, 0, st>>>(nullptr, 0, nullptr, nullptr, nullptr, (const Top2*)ws.part, 0, 0, mb);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_match(sb_point* d_pts1, int n1, const float* d_f1, const sb_point* d_pts2, int n2, const float* d_f2,
                         int nfeatures, MatchScratch& ws, int sm_count, cudaStream_t st) {
    if (n1 <= 0) return cudaSuccess;
    if (nfeatures == 64) return run_match<64>(d_pts1, n1, d_f1, d_pts2, n2, d_f2, ws, sm_count, st);
    if (nfeatures == 128) return run_match<128>(d_pts1, n1, d_f1, d_pts2, n2, d_f2, ws, sm_count, st);
    if (nfeatures < 1 || nfeatures > 256) return cudaErrorInvalidValue;
    match_generic_kernel<<<(n1 + 3) / 4, 128, 4 * nfeatures * sizeof(float), st>>>(d_pts1, n1, d_f1, d_pts2, n2 - (n2 & 31), d_f2, nfeatures);
    return cudaGetLastError();
}

cudaError_t launch_match_batch(sb_point* d_pts, long long pts_stride, const int* d_counts, const float* d_desc, long long desc_stride,
                               int npairs, const int* d_pairs, int bound, int nfeatures, MatchScratch& ws, int sm_count, cudaStream_t st) {
    if (npairs <= 0 || bound <= 0) return cudaSuccess;
    const int2* pr = reinterpret_cast<const int2*>(d_pairs);
    if (nfeatures == 64) return run_match_batch<64>(d_pts, pts_stride, d_counts, d_desc, desc_stride, npairs, pr, bound, ws, sm_count, st);
    if (nfeatures == 128) return run_match_batch<128>(d_pts, pts_stride, d_counts, d_desc, desc_stride, npairs, pr, bound, ws, sm_count, st);
    return cudaErrorInvalidValue;  // other descriptor sizes: the single-pair call (match_generic_kernel)
}

void free_match_scratch(MatchScratch& ws) {
    cudaFree(ws.a); cudaFree(ws.b); cudaFree(ws.part);
    ws = MatchScratch{};
}

}  // namespace sb
