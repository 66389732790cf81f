#!/bin/bash
# multi-GPU: raw pinned-copy ceiling at N = 1, 2, 4, 8 ranks and the bench at the box's full width (run with gpurun --gpus 8)
T=${1:-r2}
NG=$(nvidia-smi -L | wc -l)
for N in 1 2 4 8; do
  [ $N -le $NG ] || continue
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/host_copy_ceiling.py > gpurun_out/${T}_ceiling_n$N.json 2> gpurun_out/${T}_ceiling_n$N.err
  echo "ceiling N=$N rc=$?"; cat gpurun_out/${T}_ceiling_n$N.json
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 20 --warmup 5 > gpurun_out/${T}_bench_n$NG.json 2> gpurun_out/${T}_bench_n$NG.err
echo "bench N=$NG rc=$?"; cat gpurun_out/${T}_bench_n$NG.json; tail -3 gpurun_out/${T}_bench_n$NG.err
lscpu | egrep "Model name|Socket|NUMA|^CPU\(s\)" > gpurun_out/${T}_lscpu.txt; nvidia-smi topo -m > gpurun_out/${T}_topo.txt 2>&1
