// synth.cpp -- `synth_v1`, the frozen synthetic "textured frame" generator used by the benchmark
// configs of BASELINE.md. Integer-exact by construction: a splitmix64 lattice hash,
// fixed-point smoothstep value noise over six octaves, plus Gaussian blobs rendered through an
// integer exp table -- so a frame depends only on (w, h, seed, shift_x, noise) and is bit-identical
// on every host. Host code; not on the measured path (frames are generated before timing).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "surfb200.h"

namespace {

inline uint64_t mix64(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL;
    z ^= z >> 27; z *= 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return z;
}
// 24-bit lattice hash
inline uint32_t h24(int64_t ix, int64_t iy, uint64_t k, uint64_t seed) {
    uint64_t z = (uint64_t)ix * 0x9E3779B97F4A7C15ULL ^ (uint64_t)iy * 0xC2B2AE3D27D4EB4FULL ^
                 k * 0x165667B19E3779F9ULL ^ seed;
    return (uint32_t)(mix64(z) >> 40);
}

constexpr int kOct = 6;
constexpr int64_t kAmp[kOct] = {4096, 2896, 2048, 1448, 1024, 724};  // round(2^(-k/2) * 4096)
constexpr int64_t kAmpSum = 12236;
constexpr int kContrast = 60;         // grey levels per unit of value noise (frozen: ~2.4 keypoints per 1000 px at thresh 4)
constexpr int kBlobPixels = 600;      // one blob per this many pixels
constexpr int kLutN = 1024;
constexpr double kLutSpan = 4.5;      // exp(-u), u in [0, 4.5)  (cut at 3 sigma)
const float kSigmas[7] = {1.5f, 2.5f, 4.f, 6.f, 10.f, 16.f, 24.f};

struct ExpLut {
    int32_t e[kLutN];
    ExpLut() { for (int i = 0; i < kLutN; i++) e[i] = (int32_t)std::llround(65536.0 * std::exp(-(i + 0.5) * kLutSpan / kLutN)); }
};
const ExpLut& lut() { static ExpLut L; return L; }

// value noise of octave k at canvas pixel (x,y): 24-bit result
inline uint32_t vnoise(int64_t x, int64_t y, int k, uint64_t seed) {
    const int sh = k + 1;
    const int64_t P = (int64_t)1 << sh, P3 = P * P * P;
    const int64_t gx = x >> sh, gy = y >> sh, fx = x & (P - 1), fy = y & (P - 1);
    const uint64_t wx = (uint64_t)(fx * fx * (3 * P - 2 * fx)), wy = (uint64_t)(fy * fy * (3 * P - 2 * fy));
    const uint64_t ux = (uint64_t)P3 - wx, uy = (uint64_t)P3 - wy;
    const uint64_t h00 = h24(gx, gy, k, seed), h10 = h24(gx + 1, gy, k, seed);
    const uint64_t h01 = h24(gx, gy + 1, k, seed), h11 = h24(gx + 1, gy + 1, k, seed);
    const uint64_t tot = h00 * ux * uy + h10 * wx * uy + h01 * ux * wy + h11 * wx * wy;
    return (uint32_t)(tot >> (6 * sh));
}

}  // namespace

extern "C" int sb_synth_frame(uint8_t* out, int w, int h, int pitch, uint64_t seed, int shift_x, int noise_amp,
                              uint64_t noise_seed) {
    if (!out || w <= 0 || h <= 0 || pitch < w) return SB_ERR_INVALID;
    std::vector<int32_t> acc((size_t)w * h);  // grey level in Q16
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const int64_t cx = (int64_t)x + shift_x + 4096, cy = (int64_t)y + 4096;  // keep lattice coords positive
            int64_t T = 0;
            for (int k = 0; k < kOct; k++) T += kAmp[k] * (int64_t)vnoise(cx, cy, k, seed);
            const int64_t dev = T - (kAmpSum << 23);
            acc[(size_t)y * w + x] = (int32_t)((128 << 16) + ((kContrast * dev) >> 20));
        }
    // blobs live on a canvas 128 px wider than the frame so that shifted views share them
    const int cw = w + 128;
    const int nblobs = (int)(((int64_t)cw * h + kBlobPixels / 2) / kBlobPixels);
    const ExpLut& L = lut();
    for (int b = 0; b < nblobs; b++) {
        const int bx = (int)(h24(b, 0, 100, seed) % (uint32_t)cw) - 64 - shift_x;
        const int by = (int)(h24(b, 1, 100, seed) % (uint32_t)h);
        const float sigma = kSigmas[h24(b, 2, 100, seed) % 7];
        const int mag = 30 + (int)(h24(b, 3, 100, seed) % 41);        // 30..70 grey levels
        const int amp_q8 = ((h24(b, 4, 100, seed) & 1) ? mag : -mag) * 256;
        const int rad = (int)std::ceil(3.f * sigma);
        const int64_t K = (int64_t)std::llround(65536.0 * (kLutN / kLutSpan) / (2.0 * sigma * sigma));
        for (int y = by - rad; y <= by + rad; y++) {
            if (y < 0 || y >= h) continue;
            for (int x = bx - rad; x <= bx + rad; x++) {
                if (x < 0 || x >= w) continue;
                const int64_t d2 = (int64_t)(x - bx) * (x - bx) + (int64_t)(y - by) * (y - by);
                const int64_t idx = (d2 * K) >> 16;
                if (idx >= kLutN) continue;
                acc[(size_t)y * w + x] += (int32_t)(((int64_t)amp_q8 * L.e[idx]) >> 8);
            }
        }
    }
    for (int y = 0; y < h; y++) {
        uint8_t* row = out + (size_t)y * pitch;
        for (int x = 0; x < w; x++) {
            int32_t v = acc[(size_t)y * w + x];
            if (noise_amp > 0) {
                const int n = (int)(h24(x, y, 200, noise_seed) % (uint32_t)(2 * noise_amp + 1)) - noise_amp;
                v += n * 65536;
            }
            int g = (v + 32768) >> 16;
            row[x] = (uint8_t)(g < 0 ? 0 : (g > 255 ? 255 : g));
        }
        if (pitch > w) std::memset(row + w, 0, (size_t)(pitch - w));
    }
    return SB_OK;
}
