#!/bin/bash
# round-2 extras after tools/gpu_round.sh: the other BASELINE configs, the matcher (single pair and batched, with the tensor-pipe
# counter), the opt-in TMA descriptor path, per-geometry-class cost of the gather descriptor kernel
T=${1:-r2f}
timeout 900 python tools/measure_configs.py > gpurun_out/${T}_configs.jsonl 2> gpurun_out/${T}_configs.err; echo "configs rc=$?"; tail -2 gpurun_out/${T}_configs.err
timeout 200 python tools/time_match.py > gpurun_out/${T}_match.txt 2>&1; timeout 200 python tools/time_match_pairs.py >> gpurun_out/${T}_match.txt 2>&1; cat gpurun_out/${T}_match.txt
SURFB200_DESCRIBE_TMA=1 timeout 200 python tools/prof_kernels.py 64 1 > gpurun_out/${T}_tma_on.txt 2>&1; timeout 200 python tools/prof_kernels.py 64 1 >> gpurun_out/${T}_tma_on.txt 2>&1; cat gpurun_out/${T}_tma_on.txt
timeout 200 python tools/time_latency.py > gpurun_out/${T}_latency.txt 2>&1; cat gpurun_out/${T}_latency.txt
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum --clock-control none --cache-control none -k regex:match -s 9 -c 3 --csv --log-file gpurun_out/${T}_matchpairs_ll.csv python tools/time_match_pairs.py > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --cache-control none -k regex:match -s 30 -c 3 --csv --log-file gpurun_out/${T}_match_ll.csv python tools/time_match.py > /dev/null 2>&1
grep -E "gpu__time|tensor_cycles" gpurun_out/${T}_matchpairs_ll.csv gpurun_out/${T}_match_ll.csv | awk -F'","' '{print $1, $(NF-2), $NF}' | cut -c1-160
