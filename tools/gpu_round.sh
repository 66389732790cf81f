#!/bin/bash
# Round measurement: parity tests, smoke, bench (both arms), ncu launch list, ncu --set full of the stage kernels.
# usage: gpu_round.sh <tag>      (outputs under gpurun_out/<tag>_*)
set -u
T=${1:-rX}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${T}_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${T}_smoke.log
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/${T}_bench_ref.json; tail -3 gpurun_out/${T}_bench_ref.err
echo "== bench ours"; timeout 900 python bench.py > gpurun_out/${T}_bench_ours.json 2> gpurun_out/${T}_bench_ours.err; rc=$?; echo "bench rc=$rc"; cat gpurun_out/${T}_bench_ours.json; tail -3 gpurun_out/${T}_bench_ours.err
if [ $rc -eq 0 ]; then
  echo "== ncu launch list (same command as the bench, fewer steps)"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_ncu_bench.log 2>&1; echo "ncu rc=$?"
  echo "== ncu full"
  timeout 300 python tools/prof_kernels.py 8 1 > gpurun_out/${T}_prof_plain.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"hessian|describe|nms_|integral" -s 14 -c 7 -f -o gpurun_out/${T}_full python tools/prof_kernels.py 8 1 > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"; cat gpurun_out/${T}_prof_plain.log; tail -3 gpurun_out/${T}_ncu_full.log
  echo "== ncu dram traffic of the dominant kernel at the bench batch (64 frames)"
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"describe_upright|hessian|nms_scan|integral" -s 12 -c 6 --csv --log-file gpurun_out/${T}_traffic.csv python tools/prof_kernels.py 64 1 > /dev/null 2>&1; echo "traffic rc=$?"; tail -3 gpurun_out/${T}_traffic.csv
fi
