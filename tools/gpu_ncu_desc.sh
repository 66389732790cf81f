#!/bin/bash
# ncu full capture of the descriptor kernels on the 8-frame workload
set -u
T=${1:-desc}
timeout 300 python tools/prof_kernels.py 8 1 > gpurun_out/${T}_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'describe' -s 4 -c 2 -f -o gpurun_out/${T} python tools/prof_kernels.py 8 1 > gpurun_out/${T}_ncu.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/${T}_plain.log; tail -3 gpurun_out/${T}_ncu.log
