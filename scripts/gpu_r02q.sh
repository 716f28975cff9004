#!/bin/bash
# Round 2, N GPUs: the bench lines C4 and C5 (weak scaling; parity gate and bitwise sharded-vs-whole check inside the run).
TAG=${1:-r02q}; N=${2:-8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L | wc -l | tee $OUT/gpus.txt
for w in C4 C5; do
  steps=30; [ $w = C5 ] && steps=3
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $steps --warmup 3 --workload $w --no-extras 2> $OUT/bench_${w}_n$N.err | tail -1 > $OUT/bench_${w}_n$N.json
  echo "$w rc=$?"; grep -v "OMP_NUM\|\*\*\*\|^$\|NCCL version" $OUT/bench_${w}_n$N.err | tail -4
  cut -c1-300 $OUT/bench_${w}_n$N.json
done
