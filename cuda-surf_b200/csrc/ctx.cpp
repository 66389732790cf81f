// ctx.cpp -- host side of libsurfb200.so: parameter derivation, buffer geometry, launch sequencing
// and the extern "C" boundary declared in include/surfb200.h.
//
// Host logic replaced (reference file:line): Surfor::init (surf.cpp:60-91), initLut (:358-371),
// allocMemory (:374-415), the octave loop of detectAndCompute (:240-294) with the per-layer
// parameter derivation of cuCalcHessianMulti (surfd.cu:2844-2865) and cuFindMaximumWithInterp
// (surfd.cu:3062-3073), the result copies (surf.cpp:302-303, 335-342, 421-427).
// The reference re-derives and re-uploads all of this per octave per frame; here it is derived
// once in sb_create into a PipeP block that every kernel receives by value, scratch is allocated
// and zeroed once (valid regions are fully rewritten each frame, so the per-frame 23 MB memsets of
// surf.cpp:345-349 are gone), and a frame is sb_info.kernels_per_frame launches (8 for the default configuration)
// with no host round trip.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "surfb200.h"

using namespace sb;

// one batch in flight on the host-buffer path: staging buffers, events and chunk schedule. THREE sets (kHostSets): while
// batch k downloads, batch k+1 computes and batch k+2 uploads.
constexpr int kHostSets = 3;
struct HostJob {
    uint8_t* d_img = nullptr;
    sb_point* d_pts = nullptr;
    float* d_desc = nullptr;
    int* d_counts = nullptr;
    int* h_counts = nullptr;  // pinned
    std::vector<cudaEvent_t> ev_in, ev_done;
    cudaEvent_t ev_end[2] = {nullptr, nullptr};
    std::vector<int> first;   // first frame of every chunk, plus nframes
    int nframes = 0;
    bool want_desc = false, active = false, submitted = false;
};

struct sb_ctx {
    sb_params prm{};
    PipeP P{};
    int device = 0, sm_count = 148;
    cudaStream_t stream = nullptr;
    // ingest / egress streams and per-chunk events of the pipelined host-buffer path (sb_detect_batch_host)
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr, stream2 = nullptr;
    // synchronous single-frame call: count and keypoints go to the host on a side branch while the descriptor kernel runs
    cudaStream_t s_side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    HostJob job[kHostSets];
    int next_set = 0;
    // scratch, `batch` frame slots each
    int* d_integral = nullptr;
    int* d_integral_ph = nullptr;  // second copy, column-phase layout (gather Hessian)
    float* d_resp = nullptr;
    int *d_colsum = nullptr, *d_rowsum = nullptr, *d_tilesum = nullptr;
    int* d_counts = nullptr;
    unsigned* d_cand = nullptr;   // NMS candidate queues, `batch` slots of cand_cap packed words
    int* d_cand_count = nullptr;
    int* d_work = nullptr;        // per-slot work counters of the descriptor kernels (re-armed by the last nms_refine block)
    unsigned* d_refine_done = nullptr;  // per-slot count of finished nms_refine blocks (wraps to zero by itself)
    int cand_cap = 0;
    // TMA descriptor path (describe_tma.cu): tensor maps over d_integral, keypoint class lists and their counters
    void* d_desc_maps = nullptr;
    int* d_cls_idx = nullptr;
    int* d_cls_cnt = nullptr;
    uint8_t* d_up = nullptr;  // doubled=true: the 2x up-sampled frames, `batch` slots of up_pitch * P.h bytes
    int up_pitch = 0;
    int* h_counts = nullptr;          // pinned
    sb_point* h_pts = nullptr;        // pinned, max_pts
    // sb_detect_and_compute replays one captured CUDA graph per (image, points, descriptor) pointer set: the seven launches,
    // the counter memset and the two result copies of a frame go down as one submission
    struct FrameGraph { const void* img; int pitch; void* pts; void* desc; int spec; cudaGraphExec_t exec; unsigned long long used; };
    std::vector<FrameGraph> graphs;
    unsigned long long graph_clock = 0;
    int graph_failures = 0;
    float* d_desc_own = nullptr;      // fresh_desc: descriptors of the synchronous call before they are copied out
    MatchScratch match_ws;
    sb_point* h_match = nullptr;      // pinned staging of sb_match's host copy
    size_t h_match_cap = 0;
    std::string err;
};

static thread_local std::string g_create_err;

static int fail(sb_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_err = msg;
    return code;
}
#define CU(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(ctx, SB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));        \
    } while (0)

namespace sb {
bool pdl_enabled() {
    // off by default: measured on the B200 (tools/time_latency.py) the single-frame p50 went from 0.125 to 0.149 ms at 1080p and
    // from 0.367 to 0.449 ms at 4K with the dependent launches on (64-frame batches: no difference)
    static const bool on = getenv("SURFB200_PDL") && atoi(getenv("SURFB200_PDL")) != 0;
    return on;
}
}  // namespace sb

static int align_up(int a, int b) { return (a % b) ? a - a % b + b : a; }

// Derive the whole pipeline description. Returns SB_OK or an error code.
static int build_pipe(const sb_params& p, PipeP& P, std::string& why) {
    if (p.width < 32 || p.height < 32) { why = "frame smaller than 32x32"; return SB_ERR_INVALID; }
    if (p.noctaves < 1 || p.noctaves > kMaxOctave) { why = "noctaves must be 1..8"; return SB_ERR_INVALID; }
    if (p.sampling_step < 1 || p.desc_wsz < 1 || p.desc_wsz > 4 || 12 % p.desc_wsz) { why = "sampling_step>=1, desc_wsz in {1,2,3,4}"; return SB_ERR_INVALID; }
    if (p.max_pts < 1 || p.batch < 1) { why = "max_pts and batch must be >= 1"; return SB_ERR_INVALID; }
    // doubled=true: the pipeline runs on the (2w-2) x (2h-2) up-sampled frame (surf.cpp:234-235, 377-378)
    const int ew = p.doubled ? 2 * p.width - 2 : p.width, eh = p.doubled ? 2 * p.height - 2 : p.height;
    if ((long long)ew * eh * 255 > 2147483647LL) { why = "frame too large for an int32 integral image"; return SB_ERR_INVALID; }
    std::memset(&P, 0, sizeof(P));
    // SurfParam, surf.cpp:66-79
    P.doubled = p.doubled ? 1 : 0;
    P.divisor = p.doubled ? 0.5f : 1.f;
    P.init_lobe = p.init_mask_size / 3;
    P.max_scale = P.init_lobe + 2;
    P.noctaves = p.noctaves;
    P.sampling = p.sampling_step + (p.doubled ? p.sampling_step : 0);
    P.thresh = p.thresh;
    P.upright = p.upright ? 1 : 0;
    P.extend = p.extend ? 1 : 0;
    P.desc_wsz = p.desc_wsz;
    P.mag_factor = 12 / p.desc_wsz;
    P.orient_size = 4 + (p.extend ? 4 : 0);
    P.nfeatures = p.desc_wsz * p.desc_wsz * P.orient_size;
    P.max_pts = p.max_pts;
    if (P.init_lobe < 1 || P.max_scale < 3 || P.max_scale > kMaxScale) { why = "init_mask_size must give 3..8 layers per octave"; return SB_ERR_INVALID; }
    // geometry, surf.cpp:377-390
    P.w = ew; P.h = eh;
    P.iw = ew + 1; P.ih = eh + 1; P.ip = align_up(P.iw, 128);
    P.istride = (long long)P.ip * (P.ih + 2);  // one zero guard row above and below
    P.band_rows = 32;
    P.nbands = (P.h + 31) / 32;
    P.nchunks = (P.iw + 255) / 256;
    int sw = (P.iw - 1) / P.sampling, sh = (P.ih - 1) / P.sampling;
    long long roff = 0;
    int hess_tiles = 0, nms_tiles = 0;
    long long nms_cells = 0;
    // octave schedule, surf.cpp:240-294 + surfd.cu:2844-2865
    int mask = P.init_lobe - 2, octave = 1, s = 0, border1 = 0;
    int borders[kMaxScale] = {0};
    for (int o = 0; o < P.noctaves; o++) {
        OctaveP& q = P.oct[o];
        if (sw < 1 || sh < 1) { why = "too many octaves for this frame size"; return SB_ERR_INVALID; }
        if (sw > 8191 || sh > 8191) { why = "response map wider than 8191 samples (the NMS candidate queue packs row and column into 13 bits each)"; return SB_ERR_INVALID; }
        q.sw = sw; q.sh = sh; q.sp = align_up(sw, 128); q.osz = q.sh * q.sp;
        q.octave = octave; q.delta = P.sampling * octave; q.resp_off = roff;
        if (o > 0) {
            border1 = ((3 * (mask + 4 * octave)) / 2) / (P.sampling * octave) + 1;
            borders[0] = borders[1] = border1;
            s = 2;
        } else {
            border1 = ((3 * (mask + 6 * octave)) / 2) / (P.sampling * octave) + 1;
        }
        q.s0 = s; q.nl = P.max_scale - s;
        const int mask0 = mask;
        for (int i = 0, ss = s; ss < P.max_scale; i++, ss++) {
            borders[ss] = border1;  // stored before the update below: the one-layer lag of SURVEY.md 2.4-3
            q.l[i] = mask0 + 2 * octave * (i + 1);
            if (ss > 2) border1 = 3 * q.l[i] / 2 / q.delta + 1;
            q.b1[i] = border1;
            float nrm = 9.f / (float)(q.l[i] * q.l[i]);
            nrm *= nrm;
            q.norm[i] = nrm;
            mask = q.l[i];
        }
        for (int k = 0; k < P.max_scale; k++) q.borders[k] = borders[k];
        q.nmb = 0;
        int mbmin = 1 << 30;
        for (int k = 1; k < P.max_scale - 1; k += 2) {
            q.mb[q.nmb] = borders[k + 1] + 1;
            if (q.mb[q.nmb] < mbmin) mbmin = q.mb[q.nmb];
            // cells of this cell layer: rows i = mb + 2 yc < sh - mb, columns likewise (surfd.cu:690-697)
            const int cwz = (sw - 2 * q.mb[q.nmb] + 1) / 2, chz = (sh - 2 * q.mb[q.nmb] + 1) / 2;
            if (cwz > 0 && chz > 0) nms_cells += (long long)cwz * chz;
            q.nmb++;
        }
        q.hess_tile0 = hess_tiles; q.hess_tx = (sw + 31) / 32; q.hess_ty = (sh + kHessRows - 1) / kHessRows;
        hess_tiles += q.hess_tx * q.hess_ty;
        const int cw = (sw - 2 * mbmin + 1) / 2, ch = (sh - 2 * mbmin + 1) / 2;  // 2x2 cells per row / column
        q.nms_tile0 = nms_tiles;
        q.nms_tx = cw > 0 ? (cw + 31) / 32 : 0;
        q.nms_ty = ch > 0 ? (ch + 7) / 8 : 0;
        if (q.nms_tx == 0 || q.nms_ty == 0) { q.nms_tx = 1; q.nms_ty = 1; }  // keep the tile table monotone
        nms_tiles += q.nmb * q.nms_tx * q.nms_ty;
        q.inv_hess_tx = 1.f / (float)q.hess_tx;
        q.inv_nms_tx = 1.f / (float)q.nms_tx;
        roff += (long long)P.max_scale * q.osz;
        octave += octave;
        sw >>= 1; sh >>= 1;
    }
    P.rstride = roff;
    P.hess_tiles = hess_tiles;
    P.nms_tiles = nms_tiles;
    P.nms_cells = (int)std::min<long long>(std::max<long long>(nms_cells, 1), 1LL << 30);
    // tables: surf.cpp:358-371 (expf on the host, as the reference) and surf.cpp:83-90
    for (int n = 0; n < 83; n++) P.lut1[n] = expf(-(n + 0.5f) / 12.5f);
    for (int n = 0; n < 40; n++) P.lut2[n] = expf(-(n + 0.5f) / 8.f);
    P.bins[0] = (float)(-3.1415926535897932384626433832795);
    for (int i = 1; i < kNBin; i++) P.bins[i] = P.bins[i - 1] + 0.08726646259971647f;
    return SB_OK;
}

extern "C" const char* sb_last_error(const sb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

extern "C" void sb_destroy(sb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_integral); cudaFree(ctx->d_integral_ph); cudaFree(ctx->d_resp); cudaFree(ctx->d_colsum); cudaFree(ctx->d_rowsum);
    cudaFree(ctx->d_tilesum); cudaFree(ctx->d_counts); cudaFree(ctx->d_up); cudaFree(ctx->d_cand); cudaFree(ctx->d_cand_count); cudaFree(ctx->d_work); cudaFree(ctx->d_refine_done); cudaFree(ctx->d_desc_own);
    cudaFree(ctx->d_desc_maps); cudaFree(ctx->d_cls_idx); cudaFree(ctx->d_cls_cnt);
    for (auto& g : ctx->graphs) cudaGraphExecDestroy(g.exec);
    free_match_scratch(ctx->match_ws);
    if (ctx->h_counts) cudaFreeHost(ctx->h_counts);
    if (ctx->h_pts) cudaFreeHost(ctx->h_pts);
    if (ctx->h_match) cudaFreeHost(ctx->h_match);
    for (HostJob& J : ctx->job) {
        cudaFree(J.d_img); cudaFree(J.d_pts); cudaFree(J.d_desc); cudaFree(J.d_counts);
        if (J.h_counts) cudaFreeHost(J.h_counts);
        for (cudaEvent_t e : J.ev_in) cudaEventDestroy(e);
        for (cudaEvent_t e : J.ev_done) cudaEventDestroy(e);
        for (cudaEvent_t e : J.ev_end) if (e) cudaEventDestroy(e);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->s_side) cudaStreamDestroy(ctx->s_side);
    if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int sb_create(sb_ctx** out, const sb_params* params) {
    if (!out || !params) return fail(nullptr, SB_ERR_INVALID, "null argument");  // (errors of sb_create go to the thread-local slot)
    *out = nullptr;
    PipeP P;
    std::string why;
    const int rc = build_pipe(*params, P, why);
    if (rc != SB_OK) return fail(nullptr, rc, why);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, SB_ERR_CUDA, "no CUDA device: libsurfb200 has no CPU fallback");
    if (params->device < 0 || params->device >= ndev) return fail(nullptr, SB_ERR_INVALID, "device ordinal out of range");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, params->device) != cudaSuccess) return fail(nullptr, SB_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10)
        return fail(nullptr, SB_ERR_CUDA, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) + "; kernels are built for sm_100a only");
    if (cudaSetDevice(params->device) != cudaSuccess) return fail(nullptr, SB_ERR_CUDA, "cudaSetDevice failed");

    sb_ctx* c = new sb_ctx;
    c->prm = *params; c->P = P; c->device = params->device; c->sm_count = prop.multiProcessorCount;
    const int B = params->batch;
    const size_t isz = sizeof(int) * (size_t)P.istride * B;
    // + one row and one float past the last slot: with sampling_step >= 11 the lagged border of the reference's move rule
    // (surfd.cu:804-808) is 1, so the quadratic fit can read row `sh` of the last layer of the last slot
    const size_t rsz = sizeof(float) * ((size_t)P.rstride * B + (size_t)P.oct[0].sp + 1);
    const size_t tsz = sizeof(int) * (size_t)P.nbands * P.nchunks * 256 * B;
    const size_t rowsz = sizeof(int) * (size_t)P.nbands * 32 * P.nchunks * B;
    const size_t ttsz = sizeof(int) * (size_t)P.nbands * P.nchunks * B;
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    // A BLOCKING stream (cudaStreamDefault): it orders itself after everything the caller has already put on the legacy
    // default stream and the default stream orders itself after it, which is the contract of the reference (all of its
    // work runs on stream 0, so a frame produced by an earlier kernel of the caller is complete before it is read, and
    // a fill of the result buffers cannot land after the results). The synchronous entry points run here; the *_async
    // ones use the caller's stream as given.
    ok(cudaStreamCreateWithFlags(&c->stream, cudaStreamDefault));
    ok(cudaStreamCreateWithFlags(&c->s_side, cudaStreamNonBlocking));
    ok(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    ok(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    ok(cudaMalloc((void**)&c->d_integral, isz));
    ok(cudaMalloc((void**)&c->d_integral_ph, isz));
    ok(cudaMalloc((void**)&c->d_resp, rsz));
    ok(cudaMalloc((void**)&c->d_colsum, tsz));
    ok(cudaMalloc((void**)&c->d_rowsum, rowsz));
    ok(cudaMalloc((void**)&c->d_tilesum, ttsz));
    ok(cudaMalloc((void**)&c->d_counts, sizeof(int) * B));
    // one queue slot per 2x2x2 NMS cell of the frame: a cell yields at most one candidate, so the queue cannot overflow
    // whatever the threshold (1.3 MB per 1080p frame; with a capacity tied to max_pts a low threshold silently dropped
    // candidates). The keypoint append itself stays bounded by max_pts, like the reference's (surf.cpp:302-303).
    c->cand_cap = P.nms_cells;
    ok(cudaMalloc((void**)&c->d_cand, sizeof(unsigned) * (size_t)c->cand_cap * B));
    ok(cudaMalloc((void**)&c->d_cand_count, sizeof(int) * B));
    ok(cudaMalloc((void**)&c->d_work, sizeof(int) * 2 * B));  // [0, B): descriptor pass, [B, 2B): orientation pass
    ok(cudaMalloc((void**)&c->d_refine_done, sizeof(unsigned) * B));
    // opt-in (environment, read at context creation): the TMA-staged descriptor kernel for step-2 keypoints. Measured
    // on the B200 it is as fast as the gather kernel on those keypoints, not faster, so the default stays off.
    const char* tma_env = getenv("SURFB200_DESCRIBE_TMA");
    if (tma_env && atoi(tma_env) != 0 && describe_tma_applies(P)) {
        ok(cudaMalloc((void**)&c->d_cls_idx, sizeof(int) * 2 * (size_t)P.max_pts * B));
        ok(cudaMalloc((void**)&c->d_cls_cnt, sizeof(int) * 4 * B));
        if (e == cudaSuccess) ok(cudaMemset(c->d_cls_cnt, 0, sizeof(int) * 4 * B));
        if (e == cudaSuccess) ok(build_describe_maps(P, c->d_integral, B, &c->d_desc_maps));
    }
    if (P.doubled) {
        c->up_pitch = align_up(P.w, 128);
        ok(cudaMalloc((void**)&c->d_up, (size_t)c->up_pitch * P.h * B));
    }
    ok(cudaMallocHost((void**)&c->h_counts, sizeof(int) * B));
    ok(cudaMallocHost((void**)&c->h_pts, sizeof(sb_point) * (size_t)P.max_pts));
    if (e == cudaSuccess) {
        // zeroed once: borders of every response layer, row/column 0, padding and guard rows of
        // the integral are never written afterwards
        ok(cudaMemsetAsync(c->d_integral, 0, isz, c->stream));
        ok(cudaMemsetAsync(c->d_integral_ph, 0, isz, c->stream));
        ok(cudaMemsetAsync(c->d_resp, 0, rsz, c->stream));
        ok(cudaMemsetAsync(c->d_counts, 0, sizeof(int) * B, c->stream));
        ok(cudaMemsetAsync(c->d_cand_count, 0, sizeof(int) * B, c->stream));
        ok(cudaMemsetAsync(c->d_work, 0, sizeof(int) * 2 * B, c->stream));
        ok(cudaMemsetAsync(c->d_refine_done, 0, sizeof(unsigned) * B, c->stream));
        ok(cudaStreamSynchronize(c->stream));
    }
    if (e != cudaSuccess) {
        g_create_err = std::string("allocation failed: ") + cudaGetErrorString(e);
        sb_destroy(c);
        return e == cudaErrorMemoryAllocation ? SB_ERR_NOMEM : SB_ERR_CUDA;
    }
    *out = c;
    return SB_OK;
}

extern "C" int sb_get_info(const sb_ctx* ctx, sb_info* info) {
    if (!ctx || !info) return SB_ERR_INVALID;
    const PipeP& P = ctx->P;
    std::memset(info, 0, sizeof(*info));
    info->max_scale = P.max_scale; info->nfeatures = P.nfeatures;
    info->iw = P.iw; info->ih = P.ih; info->ipitch = P.ip;
    long long t = 0;
    for (int o = 0; o < P.noctaves; o++) {
        info->sw[o] = P.oct[o].sw; info->sh[o] = P.oct[o].sh; info->sp[o] = P.oct[o].sp;
        t += (long long)P.max_scale * P.oct[o].sw * P.oct[o].sh;
    }
    info->resp_floats = t;
    // (the octave-0 shared-memory Hessian kernel exists for the reference's default geometry only, hessian.cu)
    const bool fast0 = (P.sampling == 2 || P.sampling == 4) && P.init_lobe == 3 && P.max_scale == 5;
    const int nhess = fast0 ? (P.noctaves > 1 ? 2 : 1) : 1;
    info->cand_capacity = ctx->cand_cap;
    info->kernels_per_frame = (P.doubled ? 1 : 0) + 2 /*integral*/ + nhess + 2 /*nms scan, refine (which also closes the frame's counters)*/ + (P.upright ? 1 : 2) +
                              (ctx->d_desc_maps ? 2 : 0) /*classify + TMA descriptor kernel*/;
    return SB_OK;
}

// Enqueue integral -> Hessian -> NMS(+refine, append) -> clamp [-> orientation] -> descriptors.
// ev: optional 5 events recorded at the stage boundaries (profiling entry point only).
// The frames use scratch slots [slot0, slot0 + nframes).
static int enqueue_frames(sb_ctx* ctx, const uint8_t* d_images, size_t image_stride, int pitch, int nframes,
                          sb_point* d_points, int* d_counts, float* d_desc, cudaStream_t st, cudaEvent_t* ev = nullptr,
                          int slot0 = 0, int early_copy_points = -1) {
    const PipeP& P = ctx->P;
    int* integral = ctx->d_integral + (size_t)slot0 * P.istride;
    int* integral_ph = ctx->d_integral_ph + (size_t)slot0 * P.istride;
    float* resp = ctx->d_resp + (size_t)slot0 * P.rstride;
    int* colsum = ctx->d_colsum + (size_t)slot0 * P.nbands * P.nchunks * 256;
    int* rowsum = ctx->d_rowsum + (size_t)slot0 * P.nbands * 32 * P.nchunks;
    int* tilesum = ctx->d_tilesum + (size_t)slot0 * P.nbands * P.nchunks;
    if (ev) CU(cudaEventRecord(ev[0], st));
    if (P.doubled) {
        // the 2x frame replaces the caller's as the input of the integral stage
        const size_t ustride = (size_t)ctx->up_pitch * P.h;
        uint8_t* up = ctx->d_up + (size_t)slot0 * ustride;
        CU(launch_upsample2x(d_images, image_stride, pitch, ctx->prm.width, ctx->prm.height, up, ustride, ctx->up_pitch, nframes, st));
        d_images = up; image_stride = ustride; pitch = ctx->up_pitch;
    }
    CU(launch_integral(P, d_images, image_stride, pitch, nframes, integral, integral_ph, colsum, rowsum, tilesum, d_counts, st));
    if (ev) CU(cudaEventRecord(ev[1], st));
    CU(launch_hessian(P, nframes, integral, integral_ph, resp, st));
    if (ev) CU(cudaEventRecord(ev[2], st));
    CU(launch_nms(P, nframes, integral, resp, d_points, d_counts, ctx->d_cand + (size_t)slot0 * ctx->cand_cap,
                  ctx->d_cand_count + slot0, ctx->cand_cap, ctx->d_refine_done + slot0, ctx->d_work + slot0,
                  ctx->d_work + ctx->prm.batch + slot0, ctx->d_cls_cnt ? ctx->d_cls_cnt + 4 * slot0 : nullptr, st));
    if (ev) CU(cudaEventRecord(ev[3], st));
    // single-frame call: as soon as the keypoints are final (before the last descriptor kernel; the orientation pass of the
    // rotated path writes `ori` first) the count and the first points leave for the host on a side stream beside that
    // kernel, and join the main stream again at the end -- inside the captured graph this is a second branch
    struct Early { sb_ctx* c; cudaStream_t st; int* d_counts; sb_point* d_points; int n; bool done; } early_arg{ctx, st, d_counts, d_points, early_copy_points, false};
    const bool early = early_copy_points >= 0 && d_desc && ctx->s_side;
    if (d_desc) {
        DescAux aux;
        aux.maps = ctx->d_desc_maps; aux.cls_idx = ctx->d_cls_idx; aux.cls_cnt = ctx->d_cls_cnt; aux.slot0 = slot0;
        if (early) {
            aux.before_last_arg = &early_arg;
            aux.before_last = [](void* a) -> cudaError_t {
                Early& E = *static_cast<Early*>(a);
                cudaError_t e = cudaEventRecord(E.c->ev_fork, E.st);
                if (e == cudaSuccess) e = cudaStreamWaitEvent(E.c->s_side, E.c->ev_fork, 0);
                if (e == cudaSuccess) e = cudaMemcpyAsync(E.c->h_counts, E.d_counts, sizeof(int), cudaMemcpyDeviceToHost, E.c->s_side);
                if (e == cudaSuccess && E.n > 0)
                    e = cudaMemcpyAsync(E.c->h_pts, E.d_points, sizeof(sb_point) * (size_t)E.n, cudaMemcpyDeviceToHost, E.c->s_side);
                if (e == cudaSuccess) e = cudaEventRecord(E.c->ev_join, E.c->s_side);
                E.done = e == cudaSuccess;
                return e;
            };
        }
        CU(launch_describe(P, nframes, integral, d_points, P.max_pts, d_counts, -1, d_desc,
                           (long long)P.max_pts * P.nfeatures, ctx->sm_count, ctx->d_work + slot0, ctx->d_work + ctx->prm.batch + slot0, aux, st));
    }
    if (early_arg.done) CU(cudaStreamWaitEvent(st, ctx->ev_join, 0));
    if (ev) CU(cudaEventRecord(ev[4], st));
    return early_arg.done ? 1 : SB_OK;  // 1: the count and the points are already on their way
}

extern "C" int sb_detect_batch_profile(sb_ctx* ctx, const uint8_t* d_images, size_t image_stride, int pitch, int nframes,
                                       sb_point* d_points, int* d_counts, float* d_desc, void* stream, float* stage_ms) {
    if (!ctx) return SB_ERR_INVALID;
    if (!d_images || !d_points || !d_counts || !stage_ms || nframes < 1 || nframes > ctx->prm.batch || pitch < ctx->prm.width)
        return fail(ctx, SB_ERR_INVALID, "sb_detect_batch_profile: bad argument");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < 5; i++) {
        const cudaError_t ce = cudaEventCreate(&ev[i]);
        if (ce != cudaSuccess) {
            for (int k = 0; k < i; k++) cudaEventDestroy(ev[k]);
            return fail(ctx, SB_ERR_CUDA, std::string("cudaEventCreate: ") + cudaGetErrorString(ce));
        }
    }
    int rc = enqueue_frames(ctx, d_images, image_stride, pitch, nframes, d_points, d_counts, d_desc, st, ev);
    if (rc == SB_OK) {
        cudaError_t e = cudaEventSynchronize(ev[4]);
        for (int i = 0; i < 4 && e == cudaSuccess; i++) e = cudaEventElapsedTime(&stage_ms[i], ev[i], ev[i + 1]);
        if (e != cudaSuccess) rc = fail(ctx, SB_ERR_CUDA, cudaGetErrorString(e));
    }
    for (int i = 0; i < 5; i++) cudaEventDestroy(ev[i]);
    return rc;
}

extern "C" int sb_detect_batch_async(sb_ctx* ctx, const uint8_t* d_images, size_t image_stride, int pitch, int nframes,
                                     sb_point* d_points, int* d_counts, float* d_desc, void* stream) {
    if (!ctx) return SB_ERR_INVALID;
    if (!d_images || !d_points || !d_counts || nframes < 1 || nframes > ctx->prm.batch || pitch < ctx->prm.width)
        return fail(ctx, SB_ERR_INVALID, "sb_detect_batch_async: bad argument (nframes must be 1..batch, pitch >= width)");
    CU(cudaSetDevice(ctx->device));
    // `stream` is used literally: NULL is the CUDA default stream, as for any CUDA API
    return enqueue_frames(ctx, d_images, image_stride, pitch, nframes, d_points, d_counts, d_desc, (cudaStream_t)stream);
}

extern "C" int sb_sync(sb_ctx* ctx) {
    if (!ctx) return SB_ERR_INVALID;
    CU(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}

extern "C" int sb_detect_and_compute(sb_ctx* ctx, const uint8_t* d_image, int w, int h, int pitch, sb_point* d_points,
                                     sb_point* h_points, int max_pts, int* num_pts, float** d_desc_addr, int want_desc) {
    if (!ctx) return SB_ERR_INVALID;
    const PipeP& P = ctx->P;
    if (!d_image || !d_points || !num_pts) return fail(ctx, SB_ERR_INVALID, "sb_detect_and_compute: null argument");
    if (w != ctx->prm.width || h != ctx->prm.height) return fail(ctx, SB_ERR_INVALID, "sb_detect_and_compute: frame size differs from the context's (create one context per size)");
    if (max_pts != P.max_pts) return fail(ctx, SB_ERR_INVALID, "sb_detect_and_compute: max_pts differs from the context's");
    if (pitch < w) return fail(ctx, SB_ERR_INVALID, "sb_detect_and_compute: pitch < width");
    CU(cudaSetDevice(ctx->device));
    float* d_desc = nullptr;
    float* fresh = nullptr;  // fresh_desc: the buffer handed to the caller; the kernels write the context's own buffer
    if (want_desc && d_desc_addr) {
        if (ctx->prm.fresh_desc) {
            // the reference's contract (surfd.cu:3262-3266): a new buffer every call, *d_desc_addr overwritten, the caller
            // frees each of them. The frame is computed into a context-owned buffer (so the captured graph is reused) and
            // the num_pts rows are copied into the fresh allocation.
            if (!ctx->d_desc_own) CU(cudaMalloc((void**)&ctx->d_desc_own, sizeof(float) * (size_t)P.max_pts * P.nfeatures));
            d_desc = ctx->d_desc_own;
        } else {
            if (!*d_desc_addr) {
                // the callee allocates and the caller cudaFree's, like the reference; a non-NULL *d_desc_addr (the buffer
                // of the previous call, main.cpp:241-245) is reused instead of leaked
                float* buf = nullptr;
                CU(cudaMalloc((void**)&buf, sizeof(float) * (size_t)P.max_pts * P.nfeatures));
                *d_desc_addr = buf;
            }
            d_desc = *d_desc_addr;
        }
    }
    cudaStream_t st = ctx->stream;
    // One host round trip instead of two: the first kSpec points travel with the count (5 k keypoints are typical at
    // 1080p, so the copy is rarely longer than needed by more than 0.2 MB); a frame with more gets the rest afterwards.
    constexpr int kSpec = 8192;
    const int spec = h_points ? std::min(P.max_pts, kSpec) : 0;
    auto enqueue_all = [&]() -> int {
        const int rc = enqueue_frames(ctx, d_image, 0, pitch, 1, d_points, ctx->d_counts, d_desc, st, nullptr, 0, spec);
        if (rc == 1) return SB_OK;  // copied on the side branch, beside the descriptor kernel
        if (rc != SB_OK) return rc;
        CU(cudaMemcpyAsync(ctx->h_counts, ctx->d_counts, sizeof(int), cudaMemcpyDeviceToHost, st));
        if (spec > 0) CU(cudaMemcpyAsync(ctx->h_pts, d_points, sizeof(sb_point) * (size_t)spec, cudaMemcpyDeviceToHost, st));
        return SB_OK;
    };
    // One captured graph per (image, points, descriptor) pointer set, at most kMaxGraphs of them, least recently used
    // evicted (a caller that passes a new descriptor or image pointer every frame then pays a capture per call, but
    // never runs out of cache or keeps stale executables alive). A pointer set whose capture failed is launched plainly;
    // three failures switch graphs off for the context.
    constexpr size_t kMaxGraphs = 16;
    bool launched = false;
    if (ctx->graph_failures < 3) {
        sb_ctx::FrameGraph* hit = nullptr;
        for (auto& g : ctx->graphs)
            if (g.img == d_image && g.pitch == pitch && g.pts == d_points && g.desc == d_desc && g.spec == spec) { hit = &g; break; }
        if (!hit) {
            cudaGraph_t graph = nullptr;
            cudaGraphExec_t exec = nullptr;
            bool ok = cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) == cudaSuccess;
            if (ok) {
                const int rc = enqueue_all();
                ok = cudaStreamEndCapture(st, &graph) == cudaSuccess && rc == SB_OK && graph != nullptr;
            }
            if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
            if (graph) cudaGraphDestroy(graph);
            if (ok) {
                if (ctx->graphs.size() >= kMaxGraphs) {
                    size_t lru = 0;
                    for (size_t k = 1; k < ctx->graphs.size(); k++)
                        if (ctx->graphs[k].used < ctx->graphs[lru].used) lru = k;
                    cudaGraphExecDestroy(ctx->graphs[lru].exec);
                    ctx->graphs.erase(ctx->graphs.begin() + lru);
                }
                ctx->graphs.push_back({d_image, pitch, d_points, d_desc, spec, exec, 0});
                hit = &ctx->graphs.back();
            } else {
                cudaGetLastError();  // clear; this call goes down as plain launches
                ctx->graph_failures++;
            }
        }
        if (hit) {
            hit->used = ++ctx->graph_clock;
            CU(cudaGraphLaunch(hit->exec, st));
            launched = true;
        }
    }
    if (!launched) {
        const int rc = enqueue_all();
        if (rc != SB_OK) return rc;
    }
    CU(cudaStreamSynchronize(st));
    const int n = ctx->h_counts[0];
    *num_pts = n;
    if (want_desc && d_desc_addr && ctx->prm.fresh_desc) {
        // surfd.cu:3262-3264: num_pts * nfeatures floats, a new allocation per call (one float when there are no keypoints)
        CU(cudaMalloc((void**)&fresh, sizeof(float) * std::max<size_t>((size_t)n * P.nfeatures, 1)));
        if (n > 0) CU(cudaMemcpyAsync(fresh, d_desc, sizeof(float) * (size_t)n * P.nfeatures, cudaMemcpyDeviceToDevice, st));
        CU(cudaStreamSynchronize(st));
        *d_desc_addr = fresh;
    }
    if (h_points && n > 0) {
        if (n > spec) {
            CU(cudaMemcpyAsync(ctx->h_pts + spec, d_points + spec, sizeof(sb_point) * (size_t)(n - spec), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
        }
        // the reference copies the first 6 (7 with orientation) 4-byte fields of each point, surf.cpp:339
        // (two loops with compile-time sizes: the copies inline to a few moves instead of 5 k memcpy calls)
        if (want_desc && !P.upright) for (int i = 0; i < n; i++) std::memcpy(&h_points[i], &ctx->h_pts[i], 7 * sizeof(float));
        else for (int i = 0; i < n; i++) std::memcpy(&h_points[i], &ctx->h_pts[i], 6 * sizeof(float));
    }
    return SB_OK;
}

// ---- host-buffer path: submit / wait over three staging sets (kHostSets)
//
// Pipelined ingest / compute / egress. A batch is cut into chunks; chunk k's frames go up on the ingest stream, its
// kernels run on a compute stream behind an event, and as soon as its keypoint counts are visible on the host the
// exact-size copies of its points and descriptors are queued on the egress stream -- so the H2D of chunk k+1, the
// kernels of chunk k and the D2H of chunk k-1 overlap (two copy engines, PCIe is full duplex). The reference does all of
// this serially per frame with blocking copies (main.cpp:211-226, surf.cpp:302-303, 335-342).
// Chunk schedule: small chunks at both ends (the first upload and the last download are not hidden behind anything),
// full chunks in between, alternating over two compute streams so that the tail of one chunk's descriptor kernel
// overlaps the head of the next chunk. With submit / wait the caller keeps up to three batches in flight (three sets of
// staging buffers and events): batch k downloads while batch k+1 computes and batch k+2 uploads.
// streaming: another batch is in flight, so this batch's first upload and the other's last download are already hidden;
// one chunk on one compute stream (the kernels are most efficient on many frames, and two interleaved streams cost 10 %).
static void chunk_schedule(int nframes, bool streaming, std::vector<int>& first) {
    const int chunk = streaming ? nframes : (nframes >= 64 ? 16 : (nframes >= 32 ? 8 : (nframes >= 8 ? 4 : 1)));
    std::vector<int> sizes, tail;
    int left = nframes;
    for (int r = 2; !streaming && r < chunk && left > 2 * chunk; r *= 2) { sizes.push_back(r); tail.push_back(r); left -= 2 * r; }
    while (left > 0) { const int c = std::min(chunk, left); sizes.push_back(c); left -= c; }
    for (int i = (int)tail.size() - 1; i >= 0; i--) sizes.push_back(tail[i]);
    first.clear();
    int f = 0;
    for (int c : sizes) { first.push_back(f); f += c; }
    first.push_back(f);
}

extern "C" int sb_submit_batch_host(sb_ctx* ctx, const uint8_t* h_images, int nframes, int want_desc, int* ticket) {
    if (!ctx) return SB_ERR_INVALID;
    const PipeP& P = ctx->P;
    if (!h_images || !ticket || nframes < 1 || nframes > ctx->prm.batch)
        return fail(ctx, SB_ERR_INVALID, "sb_submit_batch_host: bad argument");
    const int set = ctx->next_set;
    HostJob& J = ctx->job[set];
    if (J.active) return fail(ctx, SB_ERR_INVALID, "sb_submit_batch_host: three batches are already in flight; call sb_wait_batch_host first");
    CU(cudaSetDevice(ctx->device));
    const int B = ctx->prm.batch;
    const int sw_ = ctx->prm.width, sh_ = ctx->prm.height;  // the caller's frame size (P.w, P.h are the 2x size if doubled)
    const size_t fbytes = (size_t)sw_ * sh_;
    const int dpitch = align_up(sw_, 128);
    const size_t dstride = (size_t)dpitch * sh_;
    if (!ctx->s_h2d) {
        CU(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
    }
    if (!J.d_img) {
        CU(cudaMalloc((void**)&J.d_img, dstride * B));
        CU(cudaMemset(J.d_img, 0, dstride * B));
        CU(cudaMalloc((void**)&J.d_pts, sizeof(sb_point) * (size_t)P.max_pts * B));
        CU(cudaMalloc((void**)&J.d_desc, sizeof(float) * (size_t)P.max_pts * P.nfeatures * B));
        CU(cudaMalloc((void**)&J.d_counts, sizeof(int) * B));
        CU(cudaMallocHost((void**)&J.h_counts, sizeof(int) * B));
        CU(cudaEventCreateWithFlags(&J.ev_end[0], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&J.ev_end[1], cudaEventDisableTiming));
    }
    bool streaming = false;
    for (int k = 0; k < kHostSets; k++) streaming |= k != set && ctx->job[k].active;
    chunk_schedule(nframes, streaming, J.first);
    const int nchunks = (int)J.first.size() - 1;
    while ((int)J.ev_in.size() < nchunks) {
        cudaEvent_t a = nullptr, b = nullptr;
        CU(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        J.ev_in.push_back(a);
        CU(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        J.ev_done.push_back(b);
    }
    // The previously submitted batch shares the scratch slots (integral, responses, candidate queues) of this one. Chunk
    // c of both runs on the same compute stream when their schedules are equal; otherwise order them explicitly.
    HostJob& O = ctx->job[(set + kHostSets - 1) % kHostSets];
    if (O.submitted && (O.nframes != nframes || O.first != J.first)) {
        for (cudaStream_t st : {ctx->stream, ctx->stream2}) {
            CU(cudaStreamWaitEvent(st, O.ev_end[0], 0));
            CU(cudaStreamWaitEvent(st, O.ev_end[1], 0));
        }
    }
    const size_t pstride = (size_t)P.max_pts, dstride_f = (size_t)P.max_pts * P.nfeatures;
    const int rc_enq = [&]() -> int {
    for (int k = 0; k < nchunks; k++) {
        const int f0 = J.first[k], nf = J.first[k + 1] - f0;
        cudaStream_t st = (k & 1) ? ctx->stream2 : ctx->stream;
        if (dpitch == sw_) {
            CU(cudaMemcpyAsync(J.d_img + f0 * dstride, h_images + f0 * fbytes, fbytes * nf, cudaMemcpyHostToDevice, ctx->s_h2d));
        } else {
            for (int f = f0; f < f0 + nf; f++)
                CU(cudaMemcpy2DAsync(J.d_img + f * dstride, dpitch, h_images + f * fbytes, sw_, sw_, sh_, cudaMemcpyHostToDevice, ctx->s_h2d));
        }
        CU(cudaEventRecord(J.ev_in[k], ctx->s_h2d));
        CU(cudaStreamWaitEvent(st, J.ev_in[k], 0));
        const int rc = enqueue_frames(ctx, J.d_img + f0 * dstride, dstride, dpitch, nf, J.d_pts + f0 * pstride, J.d_counts + f0,
                                      want_desc ? J.d_desc + f0 * dstride_f : nullptr, st, nullptr, f0);
        if (rc != SB_OK) return rc;
        // counts land in pinned memory of the set (the caller's array may be pageable, which would block here)
        CU(cudaMemcpyAsync(J.h_counts + f0, J.d_counts + f0, sizeof(int) * nf, cudaMemcpyDeviceToHost, st));
        CU(cudaEventRecord(J.ev_done[k], st));
    }
    CU(cudaEventRecord(J.ev_end[0], ctx->stream));
    CU(cudaEventRecord(J.ev_end[1], ctx->stream2));
    return SB_OK;
    }();
    if (rc_enq != SB_OK) {
        // some chunks may already be enqueued: let them finish, then forget this batch (its schedule must not be taken for
        // the "previous batch" of the next submit, and its staging set is free again)
        cudaStreamSynchronize(ctx->s_h2d); cudaStreamSynchronize(ctx->stream); cudaStreamSynchronize(ctx->stream2);
        J.submitted = false; J.active = false; J.nframes = 0; J.first.clear();
        return rc_enq;
    }
    J.nframes = nframes; J.want_desc = want_desc != 0; J.active = true; J.submitted = true;
    *ticket = set;
    ctx->next_set = (set + 1) % kHostSets;
    return SB_OK;
}

extern "C" int sb_wait_batch_host(sb_ctx* ctx, int ticket, sb_point* h_points, int* h_counts, float* h_desc) {
    if (!ctx) return SB_ERR_INVALID;
    if (ticket < 0 || ticket >= kHostSets || !ctx->job[ticket].active || !h_points || !h_counts)
        return fail(ctx, SB_ERR_INVALID, "sb_wait_batch_host: bad ticket or argument");
    HostJob& J = ctx->job[ticket];
    if (h_desc && !J.want_desc) return fail(ctx, SB_ERR_INVALID, "sb_wait_batch_host: the batch was submitted without descriptors");
    const PipeP& P = ctx->P;
    CU(cudaSetDevice(ctx->device));
    const size_t pstride = (size_t)P.max_pts, dstride_f = (size_t)P.max_pts * P.nfeatures;
    const int nchunks = (int)J.first.size() - 1;
    // the set is handed back only when its kernels and copies are known to be complete: a failing call below drains the
    // streams first (WAITFAIL), so a later submit cannot reuse staging buffers that are still being written
    struct Guard {
        sb_ctx* c; HostJob& j; bool ok = false;
        ~Guard() {
            if (!ok) { cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->stream2); cudaStreamSynchronize(c->s_d2h); }
            j.active = false;
        }
    } guard{ctx, J};
    for (int k = 0; k < nchunks; k++) {
        const int f0 = J.first[k], nf = J.first[k + 1] - f0;
        CU(cudaEventSynchronize(J.ev_done[k]));  // counts of chunk k are on the host; later chunks keep running
        CU(cudaStreamWaitEvent(ctx->s_d2h, J.ev_done[k], 0));
        // Only the keypoints that exist travel: one strided copy per array and chunk, as wide as the chunk's largest
        // count (a copy per frame -- 128 per 64-frame batch -- spent more time on DMA set-up than on data). Entries past a
        // frame's own count, up to that width, are overwritten with unspecified values.
        int nmax = 0;
        for (int f = f0; f < f0 + nf; f++) {
            const int n = h_counts[f] = J.h_counts[f];
            nmax = n > nmax ? n : nmax;
        }
        if (nmax <= 0) continue;
        CU(cudaMemcpy2DAsync(h_points + f0 * pstride, sizeof(sb_point) * pstride, J.d_pts + f0 * pstride, sizeof(sb_point) * pstride,
                             sizeof(sb_point) * (size_t)nmax, nf, cudaMemcpyDeviceToHost, ctx->s_d2h));
        if (h_desc)
            CU(cudaMemcpy2DAsync(h_desc + f0 * dstride_f, sizeof(float) * dstride_f, J.d_desc + f0 * dstride_f, sizeof(float) * dstride_f,
                                 sizeof(float) * (size_t)nmax * P.nfeatures, nf, cudaMemcpyDeviceToHost, ctx->s_d2h));
    }
    CU(cudaStreamSynchronize(ctx->s_d2h));
    guard.ok = true;
    return SB_OK;
}

extern "C" int sb_detect_batch_host(sb_ctx* ctx, const uint8_t* h_images, int nframes, sb_point* h_points, int* h_counts,
                                    float* h_desc) {
    if (!ctx) return SB_ERR_INVALID;
    if (!h_points || !h_counts) return fail(ctx, SB_ERR_INVALID, "sb_detect_batch_host: bad argument");
    for (const HostJob& J : ctx->job)
        if (J.active) return fail(ctx, SB_ERR_INVALID, "sb_detect_batch_host: a submitted batch is still in flight");
    int ticket = -1;
    const int rc = sb_submit_batch_host(ctx, h_images, nframes, h_desc != nullptr, &ticket);
    if (rc != SB_OK) return rc;
    return sb_wait_batch_host(ctx, ticket, h_points, h_counts, h_desc);
}

static int match_args_ok(sb_ctx* ctx, sb_point* d_pts1, int n1, const float* d_feat1, const sb_point* d_pts2, int n2,
                         const float* d_feat2) {
    if (n1 < 0 || n2 < 0 || (n1 > 0 && (!d_pts1 || !d_feat1)) || (n2 > 0 && (!d_pts2 || !d_feat2)))
        return fail(ctx, SB_ERR_INVALID, "sb_match: bad argument");
    return SB_OK;
}

extern "C" int sb_match_async(sb_ctx* ctx, sb_point* d_pts1, int n1, const float* d_feat1, const sb_point* d_pts2, int n2,
                              const float* d_feat2, void* stream) {
    if (!ctx) return SB_ERR_INVALID;
    const int rc = match_args_ok(ctx, d_pts1, n1, d_feat1, d_pts2, n2, d_feat2);
    if (rc != SB_OK || n1 == 0) return rc;
    CU(cudaSetDevice(ctx->device));
    CU(launch_match(d_pts1, n1, d_feat1, d_pts2, n2, d_feat2, ctx->P.nfeatures, ctx->match_ws, ctx->sm_count, (cudaStream_t)stream));
    return SB_OK;
}

extern "C" int sb_match_pairs_async(sb_ctx* ctx, sb_point* d_points, long long pts_stride, const int* d_counts, const float* d_desc,
                                    long long desc_stride, int npairs, const int* d_pairs, int bound, void* stream) {
    if (!ctx) return SB_ERR_INVALID;
    if (npairs < 0 || bound < 0 || (npairs > 0 && (!d_points || !d_counts || !d_desc)) || pts_stride < bound ||
        desc_stride < (long long)bound * ctx->P.nfeatures)
        return fail(ctx, SB_ERR_INVALID, "sb_match_pairs_async: bad argument");
    if (ctx->P.nfeatures != 64 && ctx->P.nfeatures != 128)
        return fail(ctx, SB_ERR_INVALID, "sb_match_pairs_async: 64- and 128-d descriptors only (use sb_match per pair)");
    if (npairs == 0 || bound == 0) return SB_OK;
    CU(cudaSetDevice(ctx->device));
    CU(launch_match_batch(d_points, pts_stride, d_counts, d_desc, desc_stride, npairs, d_pairs, bound, ctx->P.nfeatures, ctx->match_ws,
                          ctx->sm_count, (cudaStream_t)stream));
    return SB_OK;
}

extern "C" int sb_match(sb_ctx* ctx, sb_point* d_pts1, sb_point* h_pts1, int n1, const float* d_feat1,
                        const sb_point* d_pts2, int n2, const float* d_feat2) {
    if (!ctx) return SB_ERR_INVALID;
    const int rc = match_args_ok(ctx, d_pts1, n1, d_feat1, d_pts2, n2, d_feat2);
    if (rc != SB_OK || n1 == 0) return rc;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CU(launch_match(d_pts1, n1, d_feat1, d_pts2, n2, d_feat2, ctx->P.nfeatures, ctx->match_ws, ctx->sm_count, st));
    if (h_pts1) {
        // The reference copies the five match fields with a strided cudaMemcpy2D of 20-byte rows
        // (surf.cpp:421-425). One contiguous copy into pinned staging + a host scatter is ~10x faster.
        if (ctx->h_match_cap < (size_t)n1) {
            if (ctx->h_match) cudaFreeHost(ctx->h_match);
            ctx->h_match = nullptr; ctx->h_match_cap = 0;
            CU(cudaMallocHost((void**)&ctx->h_match, sizeof(sb_point) * (size_t)n1));
            ctx->h_match_cap = n1;
        }
        CU(cudaMemcpyAsync(ctx->h_match, d_pts1, sizeof(sb_point) * (size_t)n1, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        const size_t off = offsetof(sb_point, score);
        for (int i = 0; i < n1; i++) std::memcpy((char*)&h_pts1[i] + off, (const char*)&ctx->h_match[i] + off, 5 * sizeof(float));
    } else {
        CU(cudaStreamSynchronize(st));
    }
    return SB_OK;
}

extern "C" int sb_match_filter(sb_ctx* ctx, const sb_point* d_pts1, int n1, const sb_point* d_pts2, int n2, float max_ambiguity,
                               int flags, sb_pair* d_pairs, sb_pair* h_pairs, int cap, int* num_pairs) {
    if (!ctx) return SB_ERR_INVALID;
    if (!num_pairs || n1 < 0 || n2 < 0 || cap < 0 || (n1 > 0 && (!d_pts1 || !d_pairs)) ||
        ((flags & (SB_FILTER_LAPLACE | SB_FILTER_CROSS)) && n1 > 0 && !d_pts2))
        return fail(ctx, SB_ERR_INVALID, "sb_match_filter: bad argument");
    *num_pairs = 0;
    if (n1 == 0 || cap == 0) return SB_OK;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CU(launch_match_filter(d_pts1, n1, d_pts2, n2, max_ambiguity, flags, d_pairs, cap, ctx->d_counts, st));
    CU(cudaMemcpyAsync(ctx->h_counts, ctx->d_counts, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const int n = ctx->h_counts[0];
    *num_pairs = n;
    if (h_pairs && n > 0) CU(cudaMemcpy(h_pairs, d_pairs, sizeof(sb_pair) * (size_t)n, cudaMemcpyDeviceToHost));
    return SB_OK;
}

extern "C" int sb_get_integral(sb_ctx* ctx, int slot, int32_t* h_out) {
    if (!ctx || !h_out || slot < 0 || slot >= ctx->prm.batch) return SB_ERR_INVALID;
    const PipeP& P = ctx->P;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy2D(h_out, sizeof(int) * P.iw, ctx->d_integral + (size_t)slot * P.istride + P.ip, sizeof(int) * P.ip,
                    sizeof(int) * P.iw, P.ih, cudaMemcpyDeviceToHost));
    return SB_OK;
}

extern "C" int sb_get_response(sb_ctx* ctx, int slot, float* h_out) {
    if (!ctx || !h_out || slot < 0 || slot >= ctx->prm.batch) return SB_ERR_INVALID;
    const PipeP& P = ctx->P;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    float* dst = h_out;
    for (int o = 0; o < P.noctaves; o++) {
        const OctaveP& q = P.oct[o];
        for (int s = 0; s < P.max_scale; s++) {
            CU(cudaMemcpy2D(dst, sizeof(float) * q.sw, ctx->d_resp + (size_t)slot * P.rstride + q.resp_off + (size_t)s * q.osz,
                            sizeof(float) * q.sp, sizeof(float) * q.sw, q.sh, cudaMemcpyDeviceToHost));
            dst += (size_t)q.sw * q.sh;
        }
    }
    return SB_OK;
}

extern "C" int sb_describe(sb_ctx* ctx, int slot, sb_point* d_points, int n, float* d_desc) {
    if (!ctx || slot < 0 || slot >= ctx->prm.batch || n < 0 || (n > 0 && (!d_points || !d_desc))) return SB_ERR_INVALID;
    if (n == 0) return SB_OK;
    const PipeP& P = ctx->P;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemsetAsync(ctx->d_work + slot, 0, sizeof(int), ctx->stream));
    CU(cudaMemsetAsync(ctx->d_work + ctx->prm.batch + slot, 0, sizeof(int), ctx->stream));
    DescAux aux;
    // (the class lists hold max_pts indices per slot: a longer caller-supplied list takes the gather kernel only)
    if (n <= P.max_pts) { aux.maps = ctx->d_desc_maps; aux.cls_idx = ctx->d_cls_idx; aux.cls_cnt = ctx->d_cls_cnt; aux.slot0 = slot; }
    if (aux.maps) CU(cudaMemsetAsync(ctx->d_cls_cnt + 4 * slot, 0, 4 * sizeof(int), ctx->stream));
    CU(launch_describe(P, 1, ctx->d_integral + (size_t)slot * P.istride, d_points, 0, nullptr, n, d_desc, 0, ctx->sm_count, ctx->d_work + slot,
                       ctx->d_work + ctx->prm.batch + slot, aux, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}
