#!/bin/bash
# usage: gpu_ncu.sh <kernel-regex> <skip> <count> <outname> [batch]
set -u
B=${5:-8}
timeout 300 python tools/prof_kernels.py $B 1 > gpurun_out/prof_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$1" -s $2 -c $3 -o gpurun_out/$4 -f python tools/prof_kernels.py $B 1 > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/prof_plain.log; tail -3 gpurun_out/ncu_full.log
