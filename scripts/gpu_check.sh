#!/bin/bash
# One GPU-box session: smoke, parity tests, bench, ncu launch list, ncu full capture of the two CP passes.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.txt 2>&1
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.log
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -60 | tee $OUT/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 2> $OUT/bench.err | tee $OUT/bench.json
tail -5 $OUT/bench.err
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'cp_|reduce_chunks' -c 80 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
tail -3 $OUT/ncu_launches.log
echo "== ncu full"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'cp_(dual|primal)_(strip_)?kernel' -s 6 -c 2 -f -o $OUT/prof_cp \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
tail -3 $OUT/ncu_full.log
ls -la $OUT
