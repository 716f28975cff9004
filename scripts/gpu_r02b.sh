#!/bin/bash
# Round 2, second GPU call: first run of the single-sweep tv tile kernel and of the new bench.py.
TAG=${1:-r02b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -15 | tee $OUT/pytest_gpu.log
echo "== tv timing (tile)"; timeout 300 python scripts/time_tv.py 2>&1 | grep "tv " | tee $OUT/tv_times.txt
echo "== tv timing (two-sweep fallback)"; PYTVB_TV_PATH=sweeps timeout 300 python scripts/time_tv.py 2>&1 | grep "tv " | tee -a $OUT/tv_times.txt
echo "== tv timing C3 / C5 / C2"; timeout 300 python scripts/time_tv.py --shape 512 1 512 512 --rt 0 2>&1 | grep "tv " | tee -a $OUT/tv_times.txt
timeout 300 python scripts/time_tv.py --shape 64 8 2048 2048 2>&1 | grep "tv " | tee -a $OUT/tv_times.txt
timeout 300 python scripts/time_tv.py --shape 20 4 100 100 2>&1 | grep "tv " | tee -a $OUT/tv_times.txt
echo "== bench"; timeout 900 python bench.py --no-extras 2> $OUT/bench.err | tee $OUT/bench.json | cut -c1-400; tail -3 $OUT/bench.err
python -c "import bench; print(bench.lib_build_id())" > $OUT/lib_hash.txt
echo "== ncu full, tile kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tv_tile' -s 4 -c 2 -f -o $OUT/prof_tv_tile python scripts/time_tv.py hybrid --reps 3 > $OUT/ncu_tv.log 2>&1
tail -2 $OUT/ncu_tv.log
echo "== memcheck (small shapes)"; timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "readme_volume or delta" 2>&1 | tail -4 | tee $OUT/memcheck.log
