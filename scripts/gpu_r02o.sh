#!/bin/bash
# Round 2 check on one GPU with the tile kernel as the default tv path: smoke, the whole -m gpu suite, tv timings, the bench lines
# C4 / C5 / C3, small-volume latencies, per-operator bench, the ncu launch list of the C4 bench and one ncu --set full of the tile kernel.
TAG=${1:-r02o}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -6 | tee $OUT/pytest_gpu.log
echo "== tv timing"; timeout 300 python scripts/time_tv.py 2>&1 | grep "tv " | tee $OUT/tv_times.txt
timeout 300 python scripts/time_tv.py hybrid upwind central --shape 512 1 512 512 --rt 0 2>&1 | grep "tv " | tee -a $OUT/tv_times.txt
timeout 300 python scripts/time_tv.py --shape 64 8 2048 2048 2>&1 | grep "tv " | tee -a $OUT/tv_times.txt
timeout 300 python scripts/time_tv.py hybrid --shape 20 4 100 100 2>&1 | grep "tv " | tee -a $OUT/tv_times.txt
PYTVB_TV_PATH=sweeps timeout 300 python scripts/time_tv.py hybrid upwind central 2>&1 | grep "tv " | tee -a $OUT/tv_times.txt
python -c "import bench; print(bench.lib_build_id())" > $OUT/lib_hash.txt
echo "== bench C4"; timeout 900 python bench.py 2> $OUT/bench_c4.err | tee $OUT/bench_c4.json | cut -c1-300; tail -3 $OUT/bench_c4.err
echo "== bench C5"; timeout 900 python bench.py --workload C5 --steps 3 2> $OUT/bench_c5.err | tee $OUT/bench_c5.json | cut -c1-400; tail -3 $OUT/bench_c5.err
echo "== bench C3"; timeout 900 python bench.py --workload C3 --steps 200 2> $OUT/bench_c3.err | tee $OUT/bench_c3.json | cut -c1-300; tail -3 $OUT/bench_c3.err
echo "== small volumes"; timeout 600 python scripts/bench_small.py --out $OUT/small.json > $OUT/small.log 2>&1; tail -3 $OUT/small.log
echo "== ncu launch list (C4 bench)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'cp_|reduce_chunks' -c 80 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-parity-gate > $OUT/ncu_launches.log 2>&1
tail -c 300 $OUT/ncu_launches.log
echo "== ncu full, tile kernel (hybrid)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tv_tile' -s 4 -c 1 -f -o $OUT/prof_tv_tile_hybrid python scripts/time_tv.py hybrid --reps 3 > $OUT/ncu_tv.log 2>&1
# (one .ncu-rep per call: gpurun brings back at most 64 MiB)
tail -2 $OUT/ncu_tv.log
echo "== per-operator bench"; timeout 600 python scripts/bench_ops.py --out $OUT/ops.json --configs C2 C3 C4 > $OUT/ops.log 2>&1; tail -3 $OUT/ops.log
