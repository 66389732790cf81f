"""Python mirror of the reference's public interface (/root/reference/surf.h) over the C-ABI.

Names, argument meaning and ownership follow the reference:
  initSurfData / freeSurfData      surf.h:10-13, surf.cpp:10-36
  Surfor.init                      surf.h:27-29, surf.cpp:60-91
  Surfor.detectAndCompute          surf.h:36,    surf.cpp:205-355
  Surfor.match                     surf.h:40,    surf.cpp:418-428
Device buffers are torch CUDA tensors (plumbing); all compute is in libsurfb200.so.
"""
import ctypes as C

import numpy as np

from . import binding as B


def iAlignUp(a, b):
    """cuda_utils.h:160-163"""
    return a - a % b + b if a % b else a


class SurfData:
    """surf::SurfData (surf_structures.h:35-41): num_pts, max_pts, h_data (host), d_data (device)."""

    def __init__(self):
        self.num_pts = 0
        self.max_pts = 0
        self.h_data = None  # numpy structured array POINT_DTYPE [max_pts]
        self.d_data = None  # torch.uint8 CUDA tensor [max_pts*48]

    def host_points(self):
        """Full device structs of the first num_pts points, copied to the host."""
        return self.d_data[: self.num_pts * 48].cpu().numpy().view(B.POINT_DTYPE).copy()


def initSurfData(max_pts, host=True, dev=True, device=0):
    import torch
    d = SurfData()
    d.max_pts = max_pts
    d.h_data = np.zeros(max_pts, B.POINT_DTYPE) if host else None
    d.d_data = torch.zeros(max_pts * 48, dtype=torch.uint8, device=f"cuda:{device}") if dev else None
    return d


def freeSurfData(data):
    data.h_data = None
    data.d_data = None
    data.num_pts = 0
    data.max_pts = 0


class Surfor:
    """surf::Surfor. One instance per (frame size, parameter set, GPU)."""

    def __init__(self):
        self._ctx = C.c_void_p(None)
        self.params = None
        self.info = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def close(self):
        if self._ctx:
            B.lib().sb_destroy(self._ctx)
            self._ctx = C.c_void_p(None)

    # Surfor::init. The keyword-only arguments are capacities the reference takes from SurfData or has
    # no notion of (keypoints per frame, frames per batched call, device ordinal).
    def init(self, noctaves, thresh=0.2, doubled=False, init_mask_size=9, sampling_step=2, upright=False, extend=False,
             desc_wsz=4, width=-1, height=-1, *, max_pts=10000, batch=1, device=0, fresh_desc=False):
        self.close()
        p = B.SbParams(noctaves, thresh, int(doubled), init_mask_size, sampling_step, int(upright), int(extend),
                       desc_wsz, width, height, max_pts, batch, device, int(fresh_desc))
        ctx = C.c_void_p(None)
        B.check(B.lib().sb_create(C.byref(ctx), C.byref(p)), None)
        self._ctx = ctx
        self.params = p
        self.info = B.SbInfo()
        B.check(B.lib().sb_get_info(self._ctx, C.byref(self.info)), self._ctx)
        self.device = device
        return self

    @property
    def nfeatures(self):
        return self.info.nfeatures

    # Surfor::detectAndCompute(image, result, whp0, desc_addr, desc).
    #   image     torch.uint8 CUDA tensor whose rows are whp0[2] bytes apart
    #   desc_out  float32 CUDA tensor [max_pts, nfeatures] standing in for *desc_addr: passed in, it is
    #             reused; None, one is allocated here (the C-ABI would cudaMalloc it itself)
    # Returns the descriptor tensor (first result.num_pts rows valid) or None when desc=False.
    def detectAndCompute(self, image, result, whp0, desc_out=None, desc=True):
        import torch
        w, h, pitch = whp0
        n = C.c_int(0)
        hp = result.h_data.ctypes.data if result.h_data is not None else None
        addr = None
        if desc:
            if self.params.fresh_desc:
                addr = C.c_void_p(None)  # the library allocates (cudaMalloc) a new buffer per call: see detect_fresh()
            else:
                if desc_out is None:
                    desc_out = torch.empty((result.max_pts, self.nfeatures), dtype=torch.float32, device=image.device)
                addr = C.c_void_p(desc_out.data_ptr())
        rc = B.lib().sb_detect_and_compute(self._ctx, image.data_ptr(), w, h, pitch, result.d_data.data_ptr(), hp,
                                           result.max_pts, C.byref(n), C.byref(addr) if desc else None, int(desc))
        B.check(rc, self._ctx)
        result.num_pts = n.value
        if desc and self.params.fresh_desc:
            return addr.value  # raw device pointer owned by the caller (cudaFree), as with the reference
        return desc_out if desc else None

    # Surfor::match(data1, data2, features1, features2)
    def match(self, data1, data2, features1, features2):
        hp = data1.h_data.ctypes.data if data1.h_data is not None else None
        rc = B.lib().sb_match(self._ctx, data1.d_data.data_ptr(), hp, data1.num_pts, features1.data_ptr(),
                              data2.d_data.data_ptr(), data2.num_pts, features2.data_ptr())
        B.check(rc, self._ctx)

    def match_async(self, data1, data2, features1, features2, stream=None):
        """match() enqueued on a torch stream (default: current), device results only, no synchronisation"""
        import torch
        st = (stream or torch.cuda.current_stream(features1.device)).cuda_stream
        rc = B.lib().sb_match_async(self._ctx, data1.d_data.data_ptr(), data1.num_pts, features1.data_ptr(),
                                    data2.d_data.data_ptr(), data2.num_pts, features2.data_ptr(), C.c_void_p(st))
        B.check(rc, self._ctx)

    def match_pairs_async(self, points, counts, desc, npairs, bound, pairs=None, stream=None):
        """Surfor::match for all stereo pairs of a detect batch at once, counts read on the device (sb_match_pairs_async):
        points uint8 CUDA [n, max_pts*48], counts int32 CUDA [n], desc float32 CUDA [n, max_pts, nf] as written by
        detect_batch; pair z is frames (2z, 2z+1) unless `pairs` (int32 CUDA [npairs, 2]) says otherwise."""
        import torch
        st = (stream or torch.cuda.current_stream(points.device)).cuda_stream
        rc = B.lib().sb_match_pairs_async(self._ctx, points.data_ptr(), points.stride(0) // 48, counts.data_ptr(), desc.data_ptr(),
                                          desc.stride(0), npairs, pairs.data_ptr() if pairs is not None else None, bound, C.c_void_p(st))
        B.check(rc, self._ctx)

    def match_filter(self, data1, data2, max_ambiguity=0.8, laplace=False, cross=False):
        """Pairs (idx1, idx2, score, ambiguity) of the rows of data1 accepted by the consumer-side ratio test
        `ambiguity < max_ambiguity` (BASELINE config 5), optionally with equal Laplacian signs and a symmetric
        cross-check (data2 must then hold the reverse match). Returns a numpy array of PAIR_DTYPE in row order."""
        import torch
        n1 = data1.num_pts
        pairs = torch.empty((max(n1, 1), 16), dtype=torch.uint8, device=data1.d_data.device)
        host = np.zeros(max(n1, 1), B.PAIR_DTYPE)
        n = C.c_int(0)
        flags = (B.FILTER_LAPLACE if laplace else 0) | (B.FILTER_CROSS if cross else 0)
        rc = B.lib().sb_match_filter(self._ctx, data1.d_data.data_ptr(), n1, data2.d_data.data_ptr(), data2.num_pts,
                                     max_ambiguity, flags, pairs.data_ptr(), host.ctypes.data, n1, C.byref(n))
        B.check(rc, self._ctx)
        return host[: n.value].copy()

    # ---- batched forms (frame loop of main.cpp:239-245 without per-frame host round trips) ----------
    def detect_batch(self, images, pitch, points, counts, desc=None, stream=None):
        """images uint8 CUDA [n, h, pitch]; points uint8 CUDA [n, max_pts*48]; counts int32 CUDA [n];
        desc float32 CUDA [n, max_pts, nfeatures] or None. Asynchronous on `stream` (a torch stream,
        default: the current one)."""
        import torch
        n = images.shape[0]
        st = (stream or torch.cuda.current_stream(images.device)).cuda_stream
        rc = B.lib().sb_detect_batch_async(self._ctx, images.data_ptr(), images.stride(0), pitch, n, points.data_ptr(),
                                           counts.data_ptr(), desc.data_ptr() if desc is not None else None,
                                           C.c_void_p(st))
        B.check(rc, self._ctx)

    def detect_batch_profile(self, images, pitch, points, counts, desc=None, stream=None):
        """detect_batch, synchronous, returning CUDA-event times (ms) of the four stages:
        integral, Hessian, NMS+refine, (orientation+)describe."""
        import torch
        ms = (C.c_float * 4)()
        st = (stream or torch.cuda.current_stream(images.device)).cuda_stream
        rc = B.lib().sb_detect_batch_profile(self._ctx, images.data_ptr(), images.stride(0), pitch, images.shape[0],
                                             points.data_ptr(), counts.data_ptr(),
                                             desc.data_ptr() if desc is not None else None, C.c_void_p(st), ms)
        B.check(rc, self._ctx)
        return [float(v) for v in ms]

    def detect_batch_host(self, frames, points, counts, desc=None):
        """Host buffers in and out (numpy or pinned torch): frames uint8 [n, h, w]; points POINT_DTYPE
        [n, max_pts]; counts int32 [n]; desc float32 [n, max_pts, nfeatures] or None. Synchronous."""
        rc = B.lib().sb_detect_batch_host(self._ctx, _hptr(frames), frames.shape[0], _hptr(points), _hptr(counts),
                                          _hptr(desc))
        B.check(rc, self._ctx)

    def submit_batch_host(self, frames, want_desc=True):
        """First half of detect_batch_host: uploads + kernels of a batch are enqueued, a ticket comes back. At most THREE
        batches may be outstanding (three staging sets); `frames` must stay alive until the ticket has been waited for."""
        t = C.c_int(-1)
        rc = B.lib().sb_submit_batch_host(self._ctx, _hptr(frames), frames.shape[0], int(want_desc), C.byref(t))
        B.check(rc, self._ctx)
        return t.value

    def wait_batch_host(self, ticket, points, counts, desc=None):
        """Second half: downloads the batch's counts, points and descriptors into the host buffers."""
        rc = B.lib().sb_wait_batch_host(self._ctx, ticket, _hptr(points), _hptr(counts), _hptr(desc))
        B.check(rc, self._ctx)

    # ---- stage access (parity tests) ------------------------------------------------------------
    def get_integral(self, slot=0):
        out = np.empty((self.info.ih, self.info.iw), np.int32)
        B.check(B.lib().sb_get_integral(self._ctx, slot, out.ctypes.data), self._ctx)
        return out

    def get_response(self, slot=0):
        out = np.empty(self.info.resp_floats, np.float32)
        B.check(B.lib().sb_get_response(self._ctx, slot, out.ctypes.data), self._ctx)
        return out

    def split_response(self, resp):
        """flat response -> list over octaves of [max_scale, sh, sw] views"""
        out, off = [], 0
        for o in range(self.params.noctaves):
            sw, sh = self.info.sw[o], self.info.sh[o]
            n = self.info.max_scale * sw * sh
            out.append(resp[off:off + n].reshape(self.info.max_scale, sh, sw))
            off += n
        return out

    def describe(self, points_dev, n, desc_dev, slot=0):
        """(orientation +) descriptors for caller-supplied device points on the slot's integral image."""
        B.check(B.lib().sb_describe(self._ctx, slot, points_dev.data_ptr(), n, desc_dev.data_ptr()), self._ctx)


def _hptr(a):
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return a.ctypes.data
