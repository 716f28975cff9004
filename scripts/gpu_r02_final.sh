#!/bin/bash
# Round-2 final evidence on one GPU.  Order matters: the ncu --set full capture of the CP passes comes first and profiles/traffic.json
# is regenerated from it ON THE BOX (same library hash), so that the bench line that follows carries roofline.traffic of this build.
TAG=${1:-r02z}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -c "import bench; print(bench.lib_build_id())" > $OUT/lib_hash.txt; cat $OUT/lib_hash.txt
echo "== ncu full, CP passes (C4 slab)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'cp_(dual|primal)_strip' -s 6 -c 2 -f -o $OUT/prof_cp python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-parity-gate > $OUT/ncu_cp.log 2>&1
tail -2 $OUT/ncu_cp.log
python scripts/ncu_traffic.py $OUT/prof_cp.ncu-rep --lib-hash "$(cat $OUT/lib_hash.txt)" --note "ncu --set full --clock-control none, C4 slab 128x4x1024x1024 f32 hybrid Nd=8, one launch each ($TAG)" --out $OUT/traffic.json > /dev/null && cp $OUT/traffic.json profiles/traffic.json
python scripts/ncu_summary.py $OUT/prof_cp.ncu-rep $OUT/cp_ncu_full.txt > /dev/null
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -4 | tee $OUT/pytest_gpu.log
echo "== bench C4"; timeout 900 python bench.py 2> $OUT/bench_c4.err | tee $OUT/bench_c4.json | cut -c1-200; tail -3 $OUT/bench_c4.err
echo "== bench --impl reference"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2> $OUT/bench_ref.err | tee $OUT/bench_ref.json | cut -c1-300
echo "== bench C5"; timeout 900 python bench.py --workload C5 --steps 3 2> $OUT/bench_c5.err | tee $OUT/bench_c5.json | cut -c1-200; tail -3 $OUT/bench_c5.err
echo "== bench C3"; timeout 900 python bench.py --workload C3 --steps 200 2> $OUT/bench_c3.err | tee $OUT/bench_c3.json | cut -c1-200; tail -3 $OUT/bench_c3.err
echo "== ncu launch lists (C4, C5)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'cp_|reduce_chunks' -c 80 --csv --log-file $OUT/launches_c4.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-parity-gate > $OUT/ncu_launches_c4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_c5.csv \
    python bench.py --workload C5 --steps 1 --warmup 1 --no-cpu-baseline --no-extras --no-parity-gate > $OUT/ncu_launches_c5.log 2>&1
tail -c 200 $OUT/ncu_launches_c5.log
echo "== tv timing"; timeout 300 python scripts/time_tv.py 2>&1 | grep "tv " | tee $OUT/tv_times.txt
timeout 300 python scripts/time_tv.py hybrid upwind central --shape 512 1 512 512 --rt 0 2>&1 | grep "tv " | tee -a $OUT/tv_times.txt
timeout 300 python scripts/time_tv.py --shape 64 8 2048 2048 2>&1 | grep "tv " | tee -a $OUT/tv_times.txt
PYTVB_TV_PATH=sweeps timeout 300 python scripts/time_tv.py hybrid upwind central 2>&1 | grep "tv " | tee -a $OUT/tv_times.txt
PYTVB_TILE_FORM=1 timeout 300 python scripts/time_tv.py hybrid upwind central 2>&1 | grep "tv " | sed "s/^/form1 /" | tee -a $OUT/tv_times.txt
echo "== small volumes"; timeout 600 python scripts/bench_small.py --out $OUT/small.json > $OUT/small.log 2>&1; tail -3 $OUT/small.log
echo "== per-operator bench"; timeout 600 python scripts/bench_ops.py --out $OUT/ops.json --configs C2 C3 C4 > $OUT/ops.log 2>&1; tail -3 $OUT/ops.log
echo "== ncu full, tile kernel (hybrid)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tv_tile' -s 4 -c 1 -f -o $OUT/prof_tv_tile python scripts/time_tv.py hybrid --reps 3 > $OUT/ncu_tv.log 2>&1
python scripts/ncu_summary.py $OUT/prof_tv_tile.ncu-rep $OUT/tv_tile_ncu_full.txt > /dev/null
echo "== compute-sanitizer (small shapes): memcheck, racecheck, synccheck over the tile kernel and the single-GPU halo-push schedule"
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_parity.py tests/test_gpu_push_single.py -m gpu -q -x -p no:cacheprovider -k "readme_volume or delta or push" 2>&1 | grep -v "^$" | tail -6 > $OUT/sanitizer_$tool.log
  tail -2 $OUT/sanitizer_$tool.log
done
rm -f $OUT/prof_cp.ncu-rep      # its summaries stay (traffic.json, cp_ncu_full.txt); gpurun brings back at most 64 MiB
du -sh $OUT
