"""cuda-surf_b200 -- host-side mirror of the reference's surf.h interface over libsurfb200.so.

The product is the C-ABI shared library built from cuda-surf_b200/csrc (hand-written sm_100a CUDA
kernels + C++ host code). This package only binds it with ctypes and mirrors the names of the
reference's public interface (/root/reference/surf.h:10-41: initSurfData, freeSurfData,
Surfor.init / detectAndCompute / match) so tests and benchmarks read like the reference's own
main.cpp flow. torch is used for device memory, streams and torch.distributed -- plumbing only.

There is no CPU fallback: importing works anywhere, but every compute call raises SurfError when
the library or a B200 is missing.
"""
from .binding import (  # noqa: F401
    LIB_PATH, POINT_DTYPE, SbInfo, SbParams, SurfError, build_library, lib, loaded_library_path, synth_frame,
)
from .surf import Surfor, SurfData, freeSurfData, iAlignUp, initSurfData  # noqa: F401
from .sharding import shard_range  # noqa: F401

__all__ = ["Surfor", "SurfData", "initSurfData", "freeSurfData", "iAlignUp", "SurfError", "POINT_DTYPE", "lib",
           "build_library", "synth_frame", "shard_range", "LIB_PATH"]
