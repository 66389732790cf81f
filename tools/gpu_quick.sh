#!/bin/bash
# quick loop: parity tests + stage timings (+ optional bench)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
timeout 300 python tools/prof_kernels.py 8 1; timeout 300 python tools/prof_kernels.py 64 1
timeout 300 python tools/time_match.py
if [ "${1:-}" = "bench" ]; then timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"; cat gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err; fi
