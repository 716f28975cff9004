#!/bin/bash
# Round 2: the two forms of the tile kernel side by side - parity through each, timings, one ncu --set full capture of form 2.
TAG=${1:-r02m}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for form in 2 1; do
echo "== pytest -m gpu, tv forced through the tile kernel, form $form"; PYTVB_TILE_FORM=$form PYTVB_TV_PATH=tile timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py tests/test_gpu_vs_reference_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -6 | tee $OUT/pytest_gpu_tile_form$form.log
echo "== tv timing, form $form"; PYTVB_TILE_FORM=$form PYTVB_TV_PATH=tile timeout 300 python scripts/time_tv.py 2>&1 | grep "tv " | sed "s/^/form$form /" | tee -a $OUT/tv_times.txt
PYTVB_TILE_FORM=$form PYTVB_TV_PATH=tile timeout 300 python scripts/time_tv.py hybrid upwind --shape 512 1 512 512 --rt 0 2>&1 | grep "tv " | sed "s/^/form$form /" | tee -a $OUT/tv_times.txt
PYTVB_TILE_FORM=$form PYTVB_TV_PATH=tile timeout 300 python scripts/time_tv.py hybrid central --shape 64 8 2048 2048 2>&1 | grep "tv " | sed "s/^/form$form /" | tee -a $OUT/tv_times.txt
PYTVB_TILE_FORM=$form PYTVB_TV_PATH=tile timeout 300 python scripts/time_tv.py hybrid --shape 20 4 100 100 2>&1 | grep "tv " | sed "s/^/form$form /" | tee -a $OUT/tv_times.txt
done
for lib in $EXTRA_LIBS; do
  PYTVB_TV_PATH=tile PYTVB_LIB_PATH=$PWD/pytv-4d_b200/csrc/$lib timeout 300 python scripts/time_tv.py hybrid upwind central 2>&1 | grep "tv " | sed "s/^/$lib /" | tee -a $OUT/tv_times.txt
done
python -c "import bench; print(bench.lib_build_id())" > $OUT/lib_hash.txt
echo "== ncu full, tile kernel"
PYTVB_TV_PATH=tile timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tv_tile' -s 4 -c 1 -f -o $OUT/prof_tv_tile python scripts/time_tv.py hybrid --reps 3 > $OUT/ncu_tv.log 2>&1
tail -2 $OUT/ncu_tv.log
