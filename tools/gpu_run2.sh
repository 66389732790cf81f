#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== bench ours"; timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "rc=$?"; cat gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
echo "== prof plain"; timeout 300 python tools/prof_kernels.py 8 1 > gpurun_out/prof_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'hessian_kernel|describe_kernel|nms_kernel|integral_scan|integral_reduce' -s 10 -c 5 -o gpurun_out/prof_r1 python tools/prof_kernels.py 8 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"; cat gpurun_out/prof_plain.log; tail -5 gpurun_out/ncu_full.log
