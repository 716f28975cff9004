#!/bin/bash
# Round-end check on one GPU: smoke, the whole -m gpu suite, the default bench line, the ncu launch list,
# ncu --set full of the TV sweeps, per-operator bench.
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -5 | tee $OUT/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py 2> $OUT/bench.err | tee $OUT/bench.json | cut -c1-300
tail -3 $OUT/bench.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'cp_|reduce_chunks' -c 80 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $OUT/ncu_launches.log 2>&1
tail -c 300 $OUT/ncu_launches.log
echo "== tv timing"; timeout 300 python scripts/time_tv.py 2>&1 | grep "tv " | tee $OUT/tv_times.txt
echo "== ncu full, tv sweeps"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tv_(norm|grad)_strip' -s 4 -c 2 -f -o $OUT/prof_tv python scripts/time_tv.py hybrid > $OUT/ncu_tv.log 2>&1
tail -2 $OUT/ncu_tv.log
echo "== per-operator bench"; timeout 600 python scripts/bench_ops.py --out $OUT/ops.json --configs C2 C3 C4 > $OUT/ops.log 2>&1; tail -3 $OUT/ops.log
