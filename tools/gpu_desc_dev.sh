#!/bin/bash
# descriptor development loop: upright parity tests, stage timings at 8 and 64 frames, per-kernel times
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x -k "oracle or bundled or full_size or batch or describe" 2>&1 | tail -5
timeout 120 python tools/prof_kernels.py 1 1; timeout 120 python tools/prof_kernels.py 8 1; timeout 200 python tools/prof_kernels.py 64 1 && bash tools/gpu_launchlist.sh ${1:-dev} "describe|classify"
