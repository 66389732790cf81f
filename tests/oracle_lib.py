"""ctypes face of oracle/liboracle.so (CPU restatement) -- TEST INFRASTRUCTURE ONLY.

Nothing under cuda-surf_b200/ may import this module; it is the checker, never the product.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")

# numpy mirror of surf::SurfPoint (/root/reference/surf_structures.h:7-31), 48 bytes
POINT_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("scale", "<f4"), ("o", "<i4"), ("strength", "<f4"), ("laplace", "<i4"),
     ("ori", "<f4"), ("score", "<f4"), ("match", "<i4"), ("match_x", "<f4"), ("match_y", "<f4"),
     ("ambiguity", "<f4")]
)
assert POINT_DTYPE.itemsize == 48


class OrParams(C.Structure):
    _fields_ = [("thresh", C.c_float), ("init_lobe", C.c_int), ("doubled", C.c_int), ("max_scale", C.c_int),
                ("noctaves", C.c_int), ("sampling", C.c_int), ("divisor", C.c_float), ("upright", C.c_int),
                ("extend", C.c_int), ("desc_wsz", C.c_int), ("mag_factor", C.c_int), ("orient_size", C.c_int),
                ("nfeatures", C.c_int)]


class OrOctave(C.Structure):
    _fields_ = [("octave", C.c_int), ("sw", C.c_int), ("sh", C.c_int), ("s0", C.c_int), ("nl", C.c_int),
                ("l", C.c_int * 8), ("delta", C.c_int * 8), ("b1", C.c_int * 8), ("norm", C.c_float * 8),
                ("borders", C.c_int * 8), ("mb", C.c_int * 8), ("nmb", C.c_int)]


def build():
    """(Re)build liboracle.so with oracle/Makefile if it is missing or stale."""
    src = os.path.join(ORACLE_DIR, "surf_oracle.c")
    if (not os.path.exists(LIB_PATH)) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], env=env)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.or_make_params.argtypes = [C.POINTER(OrParams), C.c_int, C.c_float] + [C.c_int] * 6
        L.or_make_schedule.argtypes = [C.POINTER(OrParams), C.c_int, C.c_int, C.POINTER(OrOctave)]
        L.or_make_schedule.restype = C.c_int
        L.or_resp_floats.argtypes = [C.POINTER(OrParams), C.POINTER(OrOctave)]
        L.or_resp_floats.restype = C.c_longlong
        L.or_integral.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
        L.or_integral_doubled.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
        L.or_integral_doubled.restype = None
        L.or_hessian.argtypes = [C.POINTER(OrParams), C.POINTER(OrOctave), vp, C.c_int, C.c_int, vp]
        L.or_find_keypoints.argtypes = [C.POINTER(OrParams), C.POINTER(OrOctave), vp, C.c_int, C.c_int, vp, vp, C.c_int]
        L.or_find_keypoints.restype = C.c_int
        L.or_orientation.argtypes = [C.POINTER(OrParams), vp, C.c_int, C.c_int, vp, C.c_int]
        L.or_describe.argtypes = [C.POINTER(OrParams), vp, C.c_int, C.c_int, vp, C.c_int, vp, C.c_int]
        L.or_match.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, C.c_int]
        L.or_detect_and_compute.argtypes = [C.POINTER(OrParams), vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp]
        L.or_detect_and_compute.restype = C.c_int
        L.or_time_frames.argtypes = [C.POINTER(OrParams), vp, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.POINTER(C.c_longlong)]
        L.or_time_frames.restype = C.c_double
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Stage-by-stage CPU oracle with the argument meaning of surf::Surfor::init
    (/root/reference/surf.h:27-29)."""

    def __init__(self, noctaves=4, thresh=4.0, doubled=False, init_mask_size=9, sampling_step=2, upright=True,
                 extend=False, desc_wsz=4):
        self.p = OrParams()
        lib().or_make_params(C.byref(self.p), noctaves, thresh, int(doubled), init_mask_size, sampling_step,
                             int(upright), int(extend), desc_wsz)

    def schedule(self, w, h):
        sched = (OrOctave * 8)()
        rc = lib().or_make_schedule(C.byref(self.p), w, h, sched)
        if rc != 0:
            raise ValueError("unsupported parameters (too many scales or octaves)")
        return sched

    def integral(self, img):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w = img.shape
        if self.p.doubled:  # integral of the 2x up-sampled frame, (2h-1) x (2w-1)
            out = np.empty((2 * h - 1, 2 * w - 1), np.int32)
            lib().or_integral_doubled(_ptr(img), w, h, w, _ptr(out))
            return out
        out = np.empty((h + 1, w + 1), np.int32)
        lib().or_integral(_ptr(img), w, h, w, _ptr(out))
        return out

    def hessian(self, integral):
        h, w = integral.shape[0] - 1, integral.shape[1] - 1
        sched = self.schedule(w, h)
        n = lib().or_resp_floats(C.byref(self.p), sched)
        resp = np.empty(n, np.float32)
        lib().or_hessian(C.byref(self.p), sched, _ptr(integral), w, h, _ptr(resp))
        return resp

    def split_resp(self, resp, w, h):
        """-> list over octaves of arrays [max_scale, sh, sw] (views)."""
        sched = self.schedule(w, h)
        out, off = [], 0
        for o in range(self.p.noctaves):
            sw, sh = sched[o].sw, sched[o].sh
            n = self.p.max_scale * sw * sh
            out.append(resp[off:off + n].reshape(self.p.max_scale, sh, sw))
            off += n
        return out

    def keypoints(self, integral, resp, max_pts=65536):
        h, w = integral.shape[0] - 1, integral.shape[1] - 1
        sched = self.schedule(w, h)
        pts = np.zeros(max_pts, POINT_DTYPE)
        n = lib().or_find_keypoints(C.byref(self.p), sched, _ptr(integral), w, h, _ptr(resp), _ptr(pts), max_pts)
        return pts[:n].copy()

    def orientation(self, integral, pts):
        h, w = integral.shape[0] - 1, integral.shape[1] - 1
        pts = pts.copy()
        lib().or_orientation(C.byref(self.p), _ptr(integral), w, h, _ptr(pts), len(pts))
        return pts

    def describe(self, integral, pts, normalise=True):
        h, w = integral.shape[0] - 1, integral.shape[1] - 1
        pts = np.ascontiguousarray(pts)
        desc = np.zeros((len(pts), self.p.nfeatures), np.float32)
        lib().or_describe(C.byref(self.p), _ptr(integral), w, h, _ptr(pts), len(pts), _ptr(desc), int(normalise))
        return desc

    def detect_and_compute(self, img, max_pts=65536, desc=True):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w = img.shape
        pts = np.zeros(max_pts, POINT_DTYPE)
        d = np.zeros((max_pts, self.p.nfeatures), np.float32) if desc else None
        n = lib().or_detect_and_compute(C.byref(self.p), _ptr(img), w, h, w, _ptr(pts), max_pts,
                                        _ptr(d) if desc else None)
        if n < 0:
            raise ValueError("unsupported parameters")
        return pts[:n].copy(), (d[:n].copy() if desc else None)

    def time_frames(self, frames, max_pts=65536, threads=1):
        """frames: [n,h,w] u8. Returns (seconds, total keypoints)."""
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        n, h, w = frames.shape
        tot = C.c_longlong(0)
        s = lib().or_time_frames(C.byref(self.p), _ptr(frames), h * w, n, w, h, w, max_pts, threads, C.byref(tot))
        return s, tot.value


def match(pts1, f1, pts2, f2):
    pts1 = pts1.copy()
    f1 = np.ascontiguousarray(f1, np.float32)
    f2 = np.ascontiguousarray(f2, np.float32)
    pts2 = np.ascontiguousarray(pts2)
    lib().or_match(_ptr(pts1), len(pts1), _ptr(f1), _ptr(pts2), len(pts2), _ptr(f2), f1.shape[1])
    return pts1
