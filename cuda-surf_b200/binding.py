"""ctypes binding of libsurfb200.so (include/surfb200.h)."""
import ctypes as C
import os
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
LIB_PATH = os.path.join(PKG_DIR, "libsurfb200.so")

SB_OK, SB_ERR_INVALID, SB_ERR_CUDA, SB_ERR_UNSUPPORTED, SB_ERR_NOMEM = 0, -1, -2, -3, -4

# numpy mirror of sb_point == surf::SurfPoint (/root/reference/surf_structures.h:7-31), 48 bytes
POINT_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("scale", "<f4"), ("o", "<i4"), ("strength", "<f4"), ("laplace", "<i4"),
     ("ori", "<f4"), ("score", "<f4"), ("match", "<i4"), ("match_x", "<f4"), ("match_y", "<f4"),
     ("ambiguity", "<f4")]
)
assert POINT_DTYPE.itemsize == 48
# sb_pair (include/surfb200.h)
PAIR_DTYPE = np.dtype([("idx1", "<i4"), ("idx2", "<i4"), ("score", "<f4"), ("ambiguity", "<f4")])
FILTER_LAPLACE, FILTER_CROSS = 1, 2


class SurfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libsurfb200 error {code}: {msg}")
        self.code = code


class SbParams(C.Structure):
    _fields_ = [("noctaves", C.c_int), ("thresh", C.c_float), ("doubled", C.c_int), ("init_mask_size", C.c_int),
                ("sampling_step", C.c_int), ("upright", C.c_int), ("extend", C.c_int), ("desc_wsz", C.c_int),
                ("width", C.c_int), ("height", C.c_int), ("max_pts", C.c_int), ("batch", C.c_int), ("device", C.c_int),
                ("fresh_desc", C.c_int)]


class SbInfo(C.Structure):
    _fields_ = [("max_scale", C.c_int), ("nfeatures", C.c_int), ("iw", C.c_int), ("ih", C.c_int), ("ipitch", C.c_int),
                ("sw", C.c_int * 8), ("sh", C.c_int * 8), ("sp", C.c_int * 8), ("resp_floats", C.c_longlong),
                ("kernels_per_frame", C.c_int), ("cand_capacity", C.c_int)]


def build_library(force=False):
    """Compile libsurfb200.so in-tree for sm_100a with the repo Makefile (nvcc cross-compiles without a GPU)."""
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", ROOT, "-j8", "all"])
    return LIB_PATH


_lib = None


def lib():
    """The loaded library; raises (loudly) when it has not been built -- there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SurfError(SB_ERR_CUDA, f"{LIB_PATH} is missing: run `make` (or __graft_entry__.build()) first; "
                            "there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp, i, f, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
        L.sb_create.argtypes = [C.POINTER(vp), C.POINTER(SbParams)]
        L.sb_destroy.argtypes = [vp]
        L.sb_destroy.restype = None
        L.sb_last_error.argtypes = [vp]
        L.sb_last_error.restype = C.c_char_p
        L.sb_get_info.argtypes = [vp, C.POINTER(SbInfo)]
        L.sb_detect_and_compute.argtypes = [vp, vp, i, i, i, vp, vp, i, C.POINTER(i), C.POINTER(vp), i]
        L.sb_match.argtypes = [vp, vp, vp, i, vp, vp, i, vp]
        L.sb_match_async.argtypes = [vp, vp, i, vp, vp, i, vp, vp]
        L.sb_match_pairs_async.argtypes = [vp, vp, C.c_longlong, vp, vp, C.c_longlong, i, vp, i, vp]
        L.sb_match_filter.argtypes = [vp, vp, i, vp, i, f, i, vp, vp, i, C.POINTER(i)]
        L.sb_detect_batch_async.argtypes = [vp, vp, sz, i, i, vp, vp, vp, vp]
        L.sb_detect_batch_host.argtypes = [vp, vp, i, vp, vp, vp]
        L.sb_submit_batch_host.argtypes = [vp, vp, i, i, C.POINTER(i)]
        L.sb_wait_batch_host.argtypes = [vp, i, vp, vp, vp]
        L.sb_detect_batch_profile.argtypes = [vp, vp, sz, i, i, vp, vp, vp, vp, vp]
        L.sb_sync.argtypes = [vp]
        L.sb_get_integral.argtypes = [vp, i, vp]
        L.sb_get_response.argtypes = [vp, i, vp]
        L.sb_describe.argtypes = [vp, i, vp, i, vp]
        L.sb_synth_frame.argtypes = [vp, i, i, i, C.c_uint64, i, i, C.c_uint64]
        for name in ("sb_create", "sb_get_info", "sb_detect_and_compute", "sb_match", "sb_match_async", "sb_match_pairs_async", "sb_match_filter", "sb_detect_batch_async",
                     "sb_detect_batch_host", "sb_submit_batch_host", "sb_wait_batch_host", "sb_detect_batch_profile", "sb_sync", "sb_get_integral", "sb_get_response", "sb_describe",
                     "sb_synth_frame"):
            getattr(L, name).restype = i
        _lib = L
    return _lib


def loaded_library_path():
    return LIB_PATH if _lib is not None else None


def check(rc, ctx=None):
    if rc != SB_OK:
        msg = lib().sb_last_error(ctx)
        raise SurfError(rc, msg.decode() if msg else "")


def synth_frame(w, h, seed, shift_x=0, noise_amp=0, noise_seed=0):
    """Deterministic `synth_v1` textured frame (host, u8 [h, w]); see csrc/synth.cpp."""
    out = np.empty((h, w), np.uint8)
    check(lib().sb_synth_frame(out.ctypes.data, w, h, w, seed, shift_x, noise_amp, noise_seed))
    return out
