"""numpy restatement of the `synth_v1` workload generator (cuda-surf_b200/csrc/synth.cpp), bit-identical
to sb_synth_frame. It exists so that the REFERENCE arm of bench.py (and any checker) can build the benchmark frames
without loading the product library; tests/test_oracle.py checks the two generators against each other."""
import numpy as np

_M1, _M2 = np.uint64(0xBF58476D1CE4E5B9), np.uint64(0x94D049BB133111EB)
_KX, _KY, _KK = np.uint64(0x9E3779B97F4A7C15), np.uint64(0xC2B2AE3D27D4EB4F), np.uint64(0x165667B19E3779F9)
_AMP = [4096, 2896, 2048, 1448, 1024, 724]
_AMPSUM = 12236
_CONTRAST, _BLOBPIX, _LUTN, _LUTSPAN = 60, 600, 1024, 4.5
_SIGMAS = np.array([1.5, 2.5, 4.0, 6.0, 10.0, 16.0, 24.0], np.float32)


def _mix64(z):
    z = z ^ (z >> np.uint64(30)); z = z * _M1
    z = z ^ (z >> np.uint64(27)); z = z * _M2
    return z ^ (z >> np.uint64(31))


def _h24(ix, iy, k, seed):
    with np.errstate(over="ignore"):
        z = (np.asarray(ix).astype(np.int64).astype(np.uint64) * _KX) ^ (np.asarray(iy).astype(np.int64).astype(np.uint64) * _KY) \
            ^ (np.uint64(k) * _KK) ^ np.uint64(seed)
        return (_mix64(z) >> np.uint64(40)).astype(np.uint64)


_EXP = np.array([int(np.rint(65536.0 * np.exp(-(i + 0.5) * _LUTSPAN / _LUTN))) for i in range(_LUTN)], np.int64)


def synth_frame(w, h, seed, shift_x=0, noise_amp=0, noise_seed=0):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    cx = np.arange(w, dtype=np.int64) + shift_x + 4096
    cy = np.arange(h, dtype=np.int64) + 4096
    T = np.zeros((h, w), np.int64)
    with np.errstate(over="ignore"):
        for k in range(6):
            sh = k + 1
            P = 1 << sh
            P3 = np.uint64(P * P * P)
            gx, gy, fx, fy = cx >> sh, cy >> sh, cx & (P - 1), cy & (P - 1)
            wx = (fx * fx * (3 * P - 2 * fx)).astype(np.uint64)[None, :]
            wy = (fy * fy * (3 * P - 2 * fy)).astype(np.uint64)[:, None]
            ux, uy = P3 - wx, P3 - wy
            # hash the lattice nodes once, then gather
            nx = np.arange(gx.min(), gx.max() + 2, dtype=np.int64)
            ny = np.arange(gy.min(), gy.max() + 2, dtype=np.int64)
            H = _h24(nx[None, :], ny[:, None], k, seed)
            jx, jy = (gx - nx[0])[None, :], (gy - ny[0])[:, None]
            tot = H[jy, jx] * ux * uy + H[jy, jx + 1] * wx * uy + H[jy + 1, jx] * ux * wy + H[jy + 1, jx + 1] * wx * wy
            T += _AMP[k] * (tot >> np.uint64(6 * sh)).astype(np.int64)
    dev = T - (_AMPSUM << 23)
    acc = ((128 << 16) + ((_CONTRAST * dev) >> 20)).astype(np.int32).astype(np.int64)
    cw = w + 128
    nblobs = (cw * h + _BLOBPIX // 2) // _BLOBPIX
    b = np.arange(nblobs)
    bxs = (_h24(b, 0, 100, seed) % np.uint64(cw)).astype(np.int64) - 64 - shift_x
    bys = (_h24(b, 1, 100, seed) % np.uint64(h)).astype(np.int64)
    sig = _SIGMAS[(_h24(b, 2, 100, seed) % np.uint64(7)).astype(np.int64)]
    mag = 30 + (_h24(b, 3, 100, seed) % np.uint64(41)).astype(np.int64)
    amp = np.where((_h24(b, 4, 100, seed) & np.uint64(1)) != 0, mag, -mag) * 256
    for i in range(nblobs):
        s = float(sig[i])
        rad = int(np.ceil(np.float32(3.0) * sig[i]))
        K = int(np.rint(65536.0 * (_LUTN / _LUTSPAN) / (2.0 * s * s)))
        x0, x1 = max(int(bxs[i]) - rad, 0), min(int(bxs[i]) + rad, w - 1)
        y0, y1 = max(int(bys[i]) - rad, 0), min(int(bys[i]) + rad, h - 1)
        if x0 > x1 or y0 > y1:
            continue
        dx = np.arange(x0, x1 + 1, dtype=np.int64) - int(bxs[i])
        dy = np.arange(y0, y1 + 1, dtype=np.int64) - int(bys[i])
        idx = ((dx[None, :] ** 2 + dy[:, None] ** 2) * K) >> 16
        ok = idx < _LUTN
        add = (int(amp[i]) * _EXP[np.minimum(idx, _LUTN - 1)]) >> 8
        # the C generator accumulates in int32 (two's complement wrap never happens for these magnitudes)
        acc[y0:y1 + 1, x0:x1 + 1] += np.where(ok, add, 0)
    v = acc
    if noise_amp > 0:
        xs, ys = np.arange(w, dtype=np.int64)[None, :], np.arange(h, dtype=np.int64)[:, None]
        n = (_h24(xs + 0 * ys, ys + 0 * xs, 200, int(noise_seed) & 0xFFFFFFFFFFFFFFFF) % np.uint64(2 * noise_amp + 1)).astype(np.int64) - noise_amp
        v = v + n * 65536
    g = (v + 32768) >> 16
    return np.clip(g, 0, 255).astype(np.uint8)
