"""ctypes face of oracle/_ref/libsurfref.so -- the UNMODIFIED reference (CUDA) behind
oracle/ref_harness.cu. TEST INFRASTRUCTURE ONLY; needs a GPU. Built by `make -C oracle ref`
in the build container (where /root/reference exists) and shipped to the GPU box as a prebuilt
file; /root/reference itself is never read at run time."""
import ctypes as C
import os

import numpy as np

from oracle_lib import POINT_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libsurfref.so")

_lib = None


def available():
    return os.path.exists(REF_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(REF_PATH)
        vp, i, f = C.c_void_p, C.c_int, C.c_float
        L.ref_device_count.restype = i
        L.ref_create.argtypes = [i, i, f, i, i, i, i, i, i, i, i]
        L.ref_create.restype = vp
        L.ref_destroy.argtypes = [vp]
        L.ref_detect.argtypes = [vp, vp, i, vp, vp, i]
        L.ref_detect.restype = i
        L.ref_match.argtypes = [vp, vp, i, vp, vp, i, vp]
        L.ref_match.restype = i
        L.ref_stages.argtypes = [vp, vp, vp, vp, vp]
        L.ref_stages.restype = C.c_longlong
        L.ref_time_detect.argtypes = [vp, vp, i, i, i, vp, i]
        L.ref_time_detect.restype = i
        L.ref_time_detect_e2e.argtypes = [vp, vp, i, i, i, vp, vp]
        L.ref_time_detect_e2e.restype = i
        L.ref_time_match.argtypes = [vp, vp, i, vp, vp, i, vp, i, i, vp]
        L.ref_time_match.restype = i
        _lib = L
    return _lib


class Reference:
    """The reference's Surfor on one frame size (arguments of Surfor::init, surf.h:27-29)."""

    def __init__(self, w, h, noctaves=4, thresh=4.0, doubled=False, init_mask_size=9, sampling_step=2, upright=True,
                 extend=False, desc_wsz=4, device=0):
        self.w, self.h = w, h
        self.noctaves = noctaves
        self.doubled = bool(doubled)
        self.max_scale = init_mask_size // 3 + 2
        self.nfeatures = desc_wsz * desc_wsz * (8 if extend else 4)
        self.hnd = lib().ref_create(device, noctaves, thresh, int(doubled), init_mask_size, sampling_step, int(upright),
                                    int(extend), desc_wsz, w, h)

    def close(self):
        if self.hnd:
            lib().ref_destroy(self.hnd)
            self.hnd = None

    def detect(self, img, max_pts=65536, desc=True):
        img = np.ascontiguousarray(img, np.uint8)
        assert img.shape == (self.h, self.w)
        pts = np.zeros(max_pts, POINT_DTYPE)
        d = np.zeros((max_pts, self.nfeatures), np.float32)
        n = lib().ref_detect(self.hnd, img.ctypes.data, max_pts, pts.ctypes.data, d.ctypes.data, int(desc))
        return pts[:n].copy(), (d[:n].copy() if desc else None)

    def stages(self, img):
        """-> (integral [h+1,w+1] int32, list over octaves of [max_scale, sh, sw] float32)"""
        img = np.ascontiguousarray(img, np.uint8)
        dims = np.zeros(2 * self.noctaves, np.int32)
        n = lib().ref_stages(self.hnd, img.ctypes.data, None, None, dims.ctypes.data)
        integral = np.zeros((2 * self.h - 1, 2 * self.w - 1) if self.doubled else (self.h + 1, self.w + 1), np.int32)
        resp = np.zeros(n, np.float32)
        lib().ref_stages(self.hnd, img.ctypes.data, integral.ctypes.data, resp.ctypes.data, dims.ctypes.data)
        out, off = [], 0
        for o in range(self.noctaves):
            sw, sh = int(dims[2 * o]), int(dims[2 * o + 1])
            m = self.max_scale * sw * sh
            out.append(resp[off:off + m].reshape(self.max_scale, sh, sw))
            off += m
        return integral, out, resp

    def match(self, pts1, f1, pts2, f2):
        assert len(pts2) >= 32, "the reference reads surf2[-1] when n2 < 32 (SURVEY.md 2.4-15)"
        pts1 = pts1.copy()
        f1 = np.ascontiguousarray(f1, np.float32)
        f2 = np.ascontiguousarray(f2, np.float32)
        pts2 = np.ascontiguousarray(pts2)
        lib().ref_match(self.hnd, pts1.ctypes.data, len(pts1), f1.ctypes.data, pts2.ctypes.data, len(pts2), f2.ctypes.data)
        return pts1

    def time_detect(self, img, max_pts, warmup, iters, host_points=True):
        img = np.ascontiguousarray(img, np.uint8)
        ms = np.zeros(iters, np.float64)
        n = lib().ref_time_detect(self.hnd, img.ctypes.data, max_pts, warmup, iters, ms.ctypes.data, int(host_points))
        return ms, n

    def time_detect_e2e(self, img, max_pts, warmup, iters):
        img = np.ascontiguousarray(img, np.uint8)
        ms = np.zeros(iters, np.float64)
        d = np.zeros((max_pts, self.nfeatures), np.float32)
        n = lib().ref_time_detect_e2e(self.hnd, img.ctypes.data, max_pts, warmup, iters, ms.ctypes.data, d.ctypes.data)
        return ms, n

    def time_match(self, pts1, f1, pts2, f2, warmup, iters):
        ms = np.zeros(iters, np.float64)
        f1 = np.ascontiguousarray(f1, np.float32)
        f2 = np.ascontiguousarray(f2, np.float32)
        lib().ref_time_match(self.hnd, pts1.ctypes.data, len(pts1), f1.ctypes.data, pts2.ctypes.data, len(pts2),
                             f2.ctypes.data, warmup, iters, ms.ctypes.data)
        return ms
