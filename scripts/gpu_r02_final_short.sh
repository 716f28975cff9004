#!/bin/bash
# The essential part of gpu_r02_final.sh for the last GPU minutes of a round: traffic.json of this build, the -m gpu suite, the default
# bench line, tv timings and the ncu capture of the tile kernel.
TAG=${1:-r02zz}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -c "import bench; print(bench.lib_build_id())" > $OUT/lib_hash.txt; cat $OUT/lib_hash.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'cp_(dual|primal)_strip' -s 6 -c 2 -f -o $OUT/prof_cp python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-parity-gate > $OUT/ncu_cp.log 2>&1
python scripts/ncu_traffic.py $OUT/prof_cp.ncu-rep --lib-hash "$(cat $OUT/lib_hash.txt)" --note "ncu --set full --clock-control none, C4 slab 128x4x1024x1024 f32 hybrid Nd=8, one launch each ($TAG)" --out $OUT/traffic.json > /dev/null && cp $OUT/traffic.json profiles/traffic.json
python scripts/ncu_summary.py $OUT/prof_cp.ncu-rep $OUT/cp_ncu_full.txt > /dev/null; rm -f $OUT/prof_cp.ncu-rep
echo "== pytest -m gpu"; timeout 300 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -3 | tee $OUT/pytest_gpu.log
echo "== bench C4"; timeout 300 python bench.py 2> $OUT/bench_c4.err | tee $OUT/bench_c4.json | cut -c1-200
echo "== tv timing"; timeout 200 python scripts/time_tv.py 2>&1 | grep "tv " | tee $OUT/tv_times.txt
timeout 100 python scripts/time_tv.py hybrid upwind central --shape 512 1 512 512 --rt 0 2>&1 | grep "tv " | tee -a $OUT/tv_times.txt
timeout 100 python scripts/time_tv.py hybrid upwind central --shape 64 8 2048 2048 2>&1 | grep "tv " | tee -a $OUT/tv_times.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'tv_tile' -s 4 -c 1 -f -o $OUT/prof_tv_tile python scripts/time_tv.py hybrid --reps 3 > $OUT/ncu_tv.log 2>&1
python scripts/ncu_summary.py $OUT/prof_tv_tile.ncu-rep $OUT/tv_tile_ncu_full.txt > /dev/null
echo "== bench C5"; timeout 200 python bench.py --workload C5 --steps 3 --no-cpu-baseline 2> $OUT/bench_c5.err | tee $OUT/bench_c5.json | cut -c1-200
