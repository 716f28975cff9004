#!/bin/bash
# Round 2, two GPUs: the multi-GPU parity tests (every halo transport, ShardedTV over peer memory), then the bench lines C4 and C5 on 2 GPUs.
TAG=${1:-r02p}; N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L | tee $OUT/gpus.txt
echo "== pytest multi-GPU"; timeout 1200 python -m pytest tests/test_gpu_multigpu.py tests/test_gpu_push_single.py -x -q --tb=short -p no:cacheprovider 2>&1 | tail -8 | tee $OUT/pytest_multigpu.log
for w in C4 C5; do
  steps=30; [ $w = C5 ] && steps=3
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $steps --warmup 3 --workload $w --no-extras 2> $OUT/bench_${w}_n$N.err | tail -1 > $OUT/bench_${w}_n$N.json
  echo "$w rc=$?"; grep -v "OMP_NUM\|\*\*\*\|^$\|NCCL version" $OUT/bench_${w}_n$N.err | tail -4
  cut -c1-500 $OUT/bench_${w}_n$N.json
done
