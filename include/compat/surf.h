// surf.h -- drop-in for /root/reference/surf.h: the same free functions and the same
// surf::Surfor public interface (surf.h:10-41), implemented over the C-ABI of libsurfb200.so.
// A translation unit that included the reference's surf.h compiles against this one unchanged and
// links with -lsurfb200 instead of surf.cpp + surfd.cu.
// Errors keep the reference's behaviour: message on stderr and exit(-1) (cuda_utils.h:18-37).
#pragma once
#include "cuda_utils.h"
#include "surf_structures.h"
#include "surfb200.h"

namespace surf {

// surf.cpp:10-21
inline void initSurfData(SurfData& data, const int max_pts, const bool host, const bool dev) {
    data.num_pts = 0;
    data.max_pts = max_pts;
    const size_t size = sizeof(SurfPoint) * (size_t)max_pts;
    data.h_data = host ? (SurfPoint*)std::malloc(size) : NULL;
    data.d_data = NULL;
    if (dev) CHECK(cudaMalloc((void**)&data.d_data, size));
}

// surf.cpp:24-36
inline void freeSurfData(SurfData& data) {
    if (data.d_data != NULL) CHECK(cudaFree(data.d_data));
    if (data.h_data != NULL) std::free(data.h_data);
    data.num_pts = 0;
    data.max_pts = 0;
}

class Surfor {
public:
    Surfor() {}
    ~Surfor() { if (ctx_) sb_destroy(ctx_); }
    Surfor(const Surfor&) = delete;
    Surfor& operator=(const Surfor&) = delete;

    // surf.h:27-29. The context itself is created on the first frame, when SurfData::max_pts is
    // known (the reference also allocates its scratch lazily, surf.cpp:395-405).
    void init(const int _noctaves, const float _thresh = 0.2f, const bool _doubled = false, const int _init_mask_size = 9,
              const int _sampling_step = 2, const bool _upright = false, const bool _extend = false, const int _desc_wsz = 4,
              const int _width = -1, const int _height = -1) {
        prm_.noctaves = _noctaves; prm_.thresh = _thresh; prm_.doubled = _doubled; prm_.init_mask_size = _init_mask_size;
        prm_.sampling_step = _sampling_step; prm_.upright = _upright; prm_.extend = _extend; prm_.desc_wsz = _desc_wsz;
        prm_.width = _width; prm_.height = _height; prm_.batch = 1;
        // Descriptor buffer ownership. Default: *desc_addr is allocated on first use and REUSED when the caller passes
        // it again (main.cpp:241-245 does; the reference leaks 199 buffers there). A caller that relies on the
        // reference's literal contract -- a fresh cudaMalloc per call, the previous pointer left alone
        // (surfd.cu:3262-3266) -- builds with -DSURFB200_FRESH_DESC=1 or runs with SURFB200_FRESH_DESC=1 in the environment.
#ifdef SURFB200_FRESH_DESC
        prm_.fresh_desc = SURFB200_FRESH_DESC;
#else
        prm_.fresh_desc = 0;
#endif
        if (const char* e = std::getenv("SURFB200_FRESH_DESC")) prm_.fresh_desc = std::atoi(e) != 0;
        CHECK(cudaGetDevice(&prm_.device));
        if (ctx_) { sb_destroy(ctx_); ctx_ = NULL; }
    }

    // surf.h:36
    void detectAndCompute(unsigned char* image, SurfData& result, int3 whp0, float** desc_addr, const bool desc = true) {
        ensure(whp0.x, whp0.y, result.max_pts);
        ok(sb_detect_and_compute(ctx_, image, whp0.x, whp0.y, whp0.z, (sb_point*)result.d_data, (sb_point*)result.h_data,
                                 result.max_pts, &result.num_pts, desc_addr, desc ? 1 : 0), "detectAndCompute");
    }

    // surf.h:40
    void match(SurfData& data1, SurfData& data2, float* features1, float* features2) {
        if (!ctx_) { std::fprintf(stderr, "surf::Surfor::match called before detectAndCompute\n"); std::exit(-1); }
        ok(sb_match(ctx_, (sb_point*)data1.d_data, (data1.h_data && data1.d_data) ? (sb_point*)data1.h_data : NULL, data1.num_pts,
                    features1, (const sb_point*)data2.d_data, data2.num_pts, features2), "match");
    }

private:
    sb_params prm_ = {};
    sb_ctx* ctx_ = NULL;

    void ensure(int w, int h, int max_pts) {
        // one context per (size, capacity); a new frame size re-creates it (the reference's
        // non-`reused` path re-allocates per call, surf.cpp:228-231)
        if (ctx_ && prm_.width == w && prm_.height == h && prm_.max_pts == max_pts) return;
        if (ctx_) { sb_destroy(ctx_); ctx_ = NULL; }
        prm_.width = w; prm_.height = h; prm_.max_pts = max_pts;
        if (sb_create(&ctx_, &prm_) != SB_OK) {
            std::fprintf(stderr, "surf::Surfor: %s\n", sb_last_error(NULL));
            std::exit(-1);
        }
    }
    void ok(int rc, const char* what) {
        if (rc == SB_OK) return;
        std::fprintf(stderr, "surf::Surfor::%s: %s\n", what, sb_last_error(ctx_));
        std::exit(-1);
    }
};

}  // namespace surf
