#!/bin/bash
# Round-end check on one GPU: smoke, the whole -m gpu suite, the default bench line, the ncu launch list.
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -15 | tee $OUT/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py 2> $OUT/bench.err | tee $OUT/bench.json | cut -c1-600
tail -3 $OUT/bench.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'cp_|reduce_chunks' -c 80 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $OUT/ncu_launches.log 2>&1
tail -2 $OUT/ncu_launches.log
