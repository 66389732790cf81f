// oracle/ref_harness.cu -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A thin extern "C" face over the UNMODIFIED reference (Accustomer/CUDA-SURF), compiled from the
// sources where they lie under /root/reference by oracle/Makefile into oracle/_ref/libsurfref.so.
// The reference is CUDA, so this library needs a GPU to run: it is the parity oracle on the B200
// box (tests -m gpu, bench.py --impl reference) and the generator of tests/golden/*.npz.
//
// Two kinds of entry point:
//   * ref_detect / ref_match / ref_time_*  : through the reference's public API only
//     (surf::Surfor::init / detectAndCompute / match, /root/reference/surf.h:17-41).
//   * ref_stages : intermediates (integral, Hessian layers) are private to Surfor
//     (/root/reference/surf.h:47-48) and cleared before return (/root/reference/surf.cpp:345-349),
//     so this entry replays the host loop of /root/reference/surf.cpp:240-294 on its own buffers,
//     calling the reference's own cu* wrappers (/root/reference/surfd.h:63,76,104,111).
//
// Oracle hygiene: the reference reads uninitialised scratch on its first call,
// so every Surfor made here is warmed up with one discarded call before results are taken.
#include "surf.h"
#include "surfd.h"

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

using namespace surf;

namespace {

struct RefHandle {
    Surfor* det = nullptr;
    int w = 0, h = 0, pitch = 0;
    int noctaves = 0, init_mask = 9, sampling = 2, desc_wsz = 4;
    float thresh = 0.f;
    bool upright = true, extend = false, doubled = false;
    unsigned char* d_img = nullptr;
    bool warmed = false;
};

void upload(RefHandle* H, const uint8_t* img) {
    CHECK(cudaMemcpy2D(H->d_img, H->pitch, img, H->w, H->w, H->h, cudaMemcpyHostToDevice));
}

void warm(RefHandle* H, int max_pts) {
    if (H->warmed) return;
    SurfData tmp;
    initSurfData(tmp, max_pts, false, true);
    float* desc = nullptr;
    int3 whp = make_int3(H->w, H->h, H->pitch);
    H->det->detectAndCompute(H->d_img, tmp, whp, &desc, true);
    if (desc) CHECK(cudaFree(desc));
    freeSurfData(tmp);
    H->warmed = true;
}

}  // namespace

extern "C" {

int ref_device_count() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

void* ref_create(int device, int noctaves, float thresh, int doubled, int init_mask_size, int sampling_step,
                 int upright, int extend, int desc_wsz, int width, int height) {
    CHECK(cudaSetDevice(device));
    RefHandle* H = new RefHandle;
    H->w = width; H->h = height; H->pitch = iAlignUp(width, 128);
    H->noctaves = noctaves; H->thresh = thresh; H->init_mask = init_mask_size; H->sampling = sampling_step;
    H->upright = upright != 0; H->extend = extend != 0; H->doubled = doubled != 0; H->desc_wsz = desc_wsz;
    H->det = new Surfor;
    H->det->init(noctaves, thresh, doubled != 0, init_mask_size, sampling_step, upright != 0, extend != 0, desc_wsz,
                 width, height);
    // two spare rows: with doubled=true integralDoubleRow0U2 reads one row past the frame (surfd.cu:172-184)
    CHECK(cudaMalloc((void**)&H->d_img, (size_t)H->pitch * (height + 2)));
    CHECK(cudaMemset(H->d_img, 0, (size_t)H->pitch * (height + 2)));
    return H;
}

void ref_destroy(void* h) {
    RefHandle* H = (RefHandle*)h;
    if (!H) return;
    delete H->det;
    if (H->d_img) cudaFree(H->d_img);
    delete H;
}

// Full detect+describe through the public API. img: tight w*h u8 on the host.
// out_pts: max_pts SurfPoint (full 48-byte structs copied back from d_data);
// out_desc: max_pts*nfeatures floats (may be NULL). Returns num_pts.
int ref_detect(void* h, const uint8_t* img, int max_pts, SurfPoint* out_pts, float* out_desc, int want_desc) {
    RefHandle* H = (RefHandle*)h;
    upload(H, img);
    warm(H, max_pts);
    SurfData data;
    initSurfData(data, max_pts, false, true);
    CHECK(cudaMemset(data.d_data, 0, sizeof(SurfPoint) * max_pts));
    float* desc = nullptr;
    int3 whp = make_int3(H->w, H->h, H->pitch);
    H->det->detectAndCompute(H->d_img, data, whp, &desc, want_desc != 0);
    int n = data.num_pts;
    if (out_pts && n > 0) CHECK(cudaMemcpy(out_pts, data.d_data, sizeof(SurfPoint) * n, cudaMemcpyDeviceToHost));
    const int nfeat = H->desc_wsz * H->desc_wsz * (H->extend ? 8 : 4);
    if (want_desc && out_desc && desc && n > 0)
        CHECK(cudaMemcpy(out_desc, desc, sizeof(float) * (size_t)n * nfeat, cudaMemcpyDeviceToHost));
    if (desc) CHECK(cudaFree(desc));
    freeSurfData(data);
    return n;
}

// Matching through Surfor::match. pts1/pts2 and desc1/desc2 are host arrays; pts1 gets
// score/match/match_x/match_y/ambiguity back (whole structs copied back).
int ref_match(void* h, SurfPoint* pts1, int n1, const float* desc1, const SurfPoint* pts2, int n2, const float* desc2) {
    RefHandle* H = (RefHandle*)h;
    const int nfeat = H->desc_wsz * H->desc_wsz * (H->extend ? 8 : 4);
    SurfData d1, d2;
    // The reference writes 32-row blocks past n1: leave head-room.
    initSurfData(d1, n1 + 64, false, true);
    initSurfData(d2, n2 + 64, false, true);
    d1.num_pts = n1; d2.num_pts = n2;
    CHECK(cudaMemset(d1.d_data, 0, sizeof(SurfPoint) * (n1 + 64)));
    CHECK(cudaMemset(d2.d_data, 0, sizeof(SurfPoint) * (n2 + 64)));
    CHECK(cudaMemcpy(d1.d_data, pts1, sizeof(SurfPoint) * n1, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(d2.d_data, pts2, sizeof(SurfPoint) * n2, cudaMemcpyHostToDevice));
    float *f1 = nullptr, *f2 = nullptr;
    CHECK(cudaMalloc((void**)&f1, sizeof(float) * (size_t)(n1 + 64) * nfeat));
    CHECK(cudaMalloc((void**)&f2, sizeof(float) * (size_t)(n2 + 64) * nfeat));
    CHECK(cudaMemset(f1, 0, sizeof(float) * (size_t)(n1 + 64) * nfeat));
    CHECK(cudaMemset(f2, 0, sizeof(float) * (size_t)(n2 + 64) * nfeat));
    CHECK(cudaMemcpy(f1, desc1, sizeof(float) * (size_t)n1 * nfeat, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(f2, desc2, sizeof(float) * (size_t)n2 * nfeat, cudaMemcpyHostToDevice));
    H->det->match(d1, d2, f1, f2);
    CHECK(cudaMemcpy(pts1, d1.d_data, sizeof(SurfPoint) * n1, cudaMemcpyDeviceToHost));
    CHECK(cudaFree(f1)); CHECK(cudaFree(f2));
    freeSurfData(d1); freeSurfData(d2);
    return 0;
}

// Stage dump. Replays /root/reference/surf.cpp:240-294 with the reference's own wrappers on
// zero-initialised buffers. out_integral: tight (w+1)*(h+1) int32. out_resp: for each octave o,
// max_scale layers of tight sw_o*sh_o floats, concatenated. out_dims: noctaves*2 ints (sw,sh).
// Returns total floats written to out_resp (or the required count when out_resp==NULL).
long long ref_stages(void* h, const uint8_t* img, int32_t* out_integral, float* out_resp, int* out_dims) {
    RefHandle* H = (RefHandle*)h;
    upload(H, img);
    const int init_lobe = H->init_mask / 3;
    const int max_scale = init_lobe + 2;
    const int noct = H->noctaves;
    const int sampling = H->doubled ? 2 * H->sampling : H->sampling;  // surf.cpp:72
    int3 whp0 = make_int3(H->w, H->h, H->pitch);
    int3 iwhp = H->doubled ? make_int3(2 * H->w - 1, 2 * H->h - 1, iAlignUp(2 * H->w - 1, 128))   // surf.cpp:377-379
                           : make_int3(H->w + 1, H->h + 1, iAlignUp(H->w + 1, 128));
    int3 swhps[MAX_OCTAVE];
    int osizes[MAX_OCTAVE];
    swhps[0] = make_int3((iwhp.x - 1) / sampling, (iwhp.y - 1) / sampling, 0);
    swhps[0].z = iAlignUp(swhps[0].x, 128);
    osizes[0] = swhps[0].y * swhps[0].z;
    long long tot = (long long)osizes[0] * max_scale;
    long long tight = (long long)swhps[0].x * swhps[0].y * max_scale;
    for (int j = 1; j < noct; j++) {
        swhps[j] = make_int3(swhps[j - 1].x >> 1, swhps[j - 1].y >> 1, 0);
        swhps[j].z = iAlignUp(swhps[j].x, 128);
        osizes[j] = swhps[j].y * swhps[j].z;
        tot += (long long)osizes[j] * max_scale;
        tight += (long long)swhps[j].x * swhps[j].y * max_scale;
    }
    if (out_dims) for (int j = 0; j < noct; j++) { out_dims[2 * j] = swhps[j].x; out_dims[2 * j + 1] = swhps[j].y; }
    if (!out_resp && !out_integral) return tight;

    int* iimage = nullptr; float* tmem = nullptr;
    // four spare rows: the doubled row kernel writes rows 2h-1 and 2h of a (2h-1)-row image (surfd.cu:180-206)
    CHECK(cudaMalloc((void**)&iimage, sizeof(int) * (size_t)iwhp.z * (iwhp.y + 4)));
    CHECK(cudaMemset(iimage, 0, sizeof(int) * (size_t)iwhp.z * (iwhp.y + 4)));
    CHECK(cudaMalloc((void**)&tmem, sizeof(float) * tot));
    CHECK(cudaMemset(tmem, 0, sizeof(float) * tot));

    if (H->doubled) cuIntegralDoubleU4(H->d_img, iimage, whp0, iwhp);
    else cuIntegral(H->d_img, iimage, whp0, iwhp);
    int mask_size = init_lobe - 2, s = 0, octave = 1, border1 = 0, offset = 0;
    int borders[MAX_OCTAVE];
    for (int o = 0; o < noct; o++) {
        if (o > 0) {
            cuHalfImage(tmem + offset - 3 * osizes[o - 1], tmem + offset, swhps[o - 1], swhps[o]);
            cuHalfImage(tmem + offset - 1 * osizes[o - 1], tmem + offset + osizes[o], swhps[o - 1], swhps[o]);
            border1 = ((3 * (mask_size + 4 * octave)) / 2) / (sampling * octave) + 1;
            borders[0] = border1; borders[1] = border1; s = 2;
        } else {
            border1 = ((3 * (mask_size + 6 * octave)) / 2) / (sampling * octave) + 1;
        }
        cuCalcHessianMulti(iimage, tmem + offset, iwhp, swhps[o], s, max_scale, mask_size, border1, borders, octave,
                           sampling);
        offset += max_scale * osizes[o];
        octave += octave;
    }
    CHECK(cudaDeviceSynchronize());
    if (out_integral)
        CHECK(cudaMemcpy2D(out_integral, sizeof(int) * iwhp.x, iimage, sizeof(int) * iwhp.z, sizeof(int) * iwhp.x, iwhp.y,
                           cudaMemcpyDeviceToHost));
    if (out_resp) {
        float* dst = out_resp; long long off = 0;
        for (int o = 0; o < noct; o++) {
            for (int l = 0; l < max_scale; l++) {
                CHECK(cudaMemcpy2D(dst, sizeof(float) * swhps[o].x, tmem + off + (long long)l * osizes[o],
                                   sizeof(float) * swhps[o].z, sizeof(float) * swhps[o].x, swhps[o].y,
                                   cudaMemcpyDeviceToHost));
                dst += (long long)swhps[o].x * swhps[o].y;
            }
            off += (long long)max_scale * osizes[o];
        }
    }
    CHECK(cudaFree(iimage)); CHECK(cudaFree(tmem));
    return tight;
}

// Timing of the reference's own path, as main.cpp:239-245 runs it: synchronous detectAndCompute
// on a device-resident frame, host wall-clock per call. The descriptor buffer the reference
// allocates per call (surfd.cu:3264) is freed outside the timed region. ms_out: iters doubles.
int ref_time_detect(void* h, const uint8_t* img, int max_pts, int warmup, int iters, double* ms_out, int host_points) {
    RefHandle* H = (RefHandle*)h;
    upload(H, img);
    warm(H, max_pts);
    SurfData data;
    initSurfData(data, max_pts, host_points != 0, true);
    int3 whp = make_int3(H->w, H->h, H->pitch);
    int n = 0;
    for (int it = 0; it < warmup + iters; it++) {
        float* desc = nullptr;
        CHECK(cudaDeviceSynchronize());
        auto t0 = std::chrono::steady_clock::now();
        H->det->detectAndCompute(H->d_img, data, whp, &desc, true);
        CHECK(cudaDeviceSynchronize());
        auto t1 = std::chrono::steady_clock::now();
        if (it >= warmup) ms_out[it - warmup] = std::chrono::duration<double, std::milli>(t1 - t0).count();
        if (desc) CHECK(cudaFree(desc));
        n = data.num_pts;
    }
    freeSurfData(data);
    return n;
}

// End-to-end variant: host frame in (H2D inside the timed region), host points + descriptors out.
int ref_time_detect_e2e(void* h, const uint8_t* img, int max_pts, int warmup, int iters, double* ms_out,
                        float* out_desc) {
    RefHandle* H = (RefHandle*)h;
    upload(H, img);
    warm(H, max_pts);
    SurfData data;
    initSurfData(data, max_pts, true, true);
    int3 whp = make_int3(H->w, H->h, H->pitch);
    const int nfeat = H->desc_wsz * H->desc_wsz * (H->extend ? 8 : 4);
    int n = 0;
    for (int it = 0; it < warmup + iters; it++) {
        float* desc = nullptr;
        CHECK(cudaDeviceSynchronize());
        auto t0 = std::chrono::steady_clock::now();
        upload(H, img);
        H->det->detectAndCompute(H->d_img, data, whp, &desc, true);
        if (out_desc && desc && data.num_pts > 0)
            CHECK(cudaMemcpy(out_desc, desc, sizeof(float) * (size_t)data.num_pts * nfeat, cudaMemcpyDeviceToHost));
        CHECK(cudaDeviceSynchronize());
        auto t1 = std::chrono::steady_clock::now();
        if (it >= warmup) ms_out[it - warmup] = std::chrono::duration<double, std::milli>(t1 - t0).count();
        if (desc) CHECK(cudaFree(desc));
        n = data.num_pts;
    }
    freeSurfData(data);
    return n;
}

int ref_time_match(void* h, const SurfPoint* pts1, int n1, const float* desc1, const SurfPoint* pts2, int n2,
                   const float* desc2, int warmup, int iters, double* ms_out) {
    RefHandle* H = (RefHandle*)h;
    const int nfeat = H->desc_wsz * H->desc_wsz * (H->extend ? 8 : 4);
    SurfData d1, d2;
    initSurfData(d1, n1 + 64, true, true);
    initSurfData(d2, n2 + 64, false, true);
    d1.num_pts = n1; d2.num_pts = n2;
    CHECK(cudaMemcpy(d1.d_data, pts1, sizeof(SurfPoint) * n1, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(d2.d_data, pts2, sizeof(SurfPoint) * n2, cudaMemcpyHostToDevice));
    float *f1 = nullptr, *f2 = nullptr;
    CHECK(cudaMalloc((void**)&f1, sizeof(float) * (size_t)(n1 + 64) * nfeat));
    CHECK(cudaMalloc((void**)&f2, sizeof(float) * (size_t)(n2 + 64) * nfeat));
    CHECK(cudaMemset(f1, 0, sizeof(float) * (size_t)(n1 + 64) * nfeat));
    CHECK(cudaMemset(f2, 0, sizeof(float) * (size_t)(n2 + 64) * nfeat));
    CHECK(cudaMemcpy(f1, desc1, sizeof(float) * (size_t)n1 * nfeat, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(f2, desc2, sizeof(float) * (size_t)n2 * nfeat, cudaMemcpyHostToDevice));
    for (int it = 0; it < warmup + iters; it++) {
        CHECK(cudaDeviceSynchronize());
        auto t0 = std::chrono::steady_clock::now();
        H->det->match(d1, d2, f1, f2);
        CHECK(cudaDeviceSynchronize());
        auto t1 = std::chrono::steady_clock::now();
        if (it >= warmup) ms_out[it - warmup] = std::chrono::duration<double, std::milli>(t1 - t0).count();
    }
    CHECK(cudaFree(f1)); CHECK(cudaFree(f2));
    freeSurfData(d1); freeSurfData(d2);
    return 0;
}

}  // extern "C"
