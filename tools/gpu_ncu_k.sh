#!/bin/bash
# usage: gpu_ncu_k.sh <tag> <kernel-regex> [skip] [count]   -- ncu full capture of selected kernels on the 8-frame workload
set -u
T=$1; K=$2; S=${3:-4}; C=${4:-3}
timeout 300 python tools/prof_kernels.py 8 ${UPR:-1} > gpurun_out/${T}_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -s $S -c $C -f -o gpurun_out/${T} python tools/prof_kernels.py 8 ${UPR:-1} > gpurun_out/${T}_ncu.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/${T}_plain.log; tail -2 gpurun_out/${T}_ncu.log
