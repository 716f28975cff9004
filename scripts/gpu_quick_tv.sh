#!/bin/bash
# Quick check of a tile-kernel change: parity tests through the tile path, timings of the usual shapes (no ncu).
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu (tv through the tile kernel)"; PYTVB_TV_PATH=tile timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py tests/test_gpu_vs_reference_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -3 | tee $OUT/pytest_gpu_tile.log
timeout 300 python scripts/time_tv.py 2>&1 | grep "tv " | tee $OUT/tv_times.txt
timeout 300 python scripts/time_tv.py hybrid upwind central --shape 512 1 512 512 --rt 0 2>&1 | grep "tv " | tee -a $OUT/tv_times.txt
timeout 300 python scripts/time_tv.py hybrid central --shape 64 8 2048 2048 2>&1 | grep "tv " | tee -a $OUT/tv_times.txt
timeout 300 python scripts/time_tv.py hybrid --shape 20 4 100 100 2>&1 | grep "tv " | tee -a $OUT/tv_times.txt
for lib in $EXTRA_LIBS; do
  PYTVB_LIB_PATH=$PWD/pytv-4d_b200/csrc/$lib timeout 300 python scripts/time_tv.py hybrid upwind central 2>&1 | grep "tv " | sed "s/^/$lib /" | tee -a $OUT/tv_times.txt
done
