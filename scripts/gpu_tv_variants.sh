#!/bin/bash
# Timings of tv_<scheme> for the default library, the two-sweep fallback and tuning variants ($EXTRA_LIBS).
OUT=gpurun_out/${1:-tvvar}; mkdir -p $OUT
timeout 300 python scripts/time_tv.py hybrid upwind central 2>&1 | grep "tv " | tee $OUT/tv_times.txt
PYTVB_TV_PATH=sweeps timeout 300 python scripts/time_tv.py hybrid upwind central 2>&1 | grep "tv " | tee -a $OUT/tv_times.txt
PYTVB_TV_PATH=sweeps timeout 300 python scripts/time_tv.py hybrid central --shape 64 8 2048 2048 2>&1 | grep "tv " | tee -a $OUT/tv_times.txt
PYTVB_TV_PATH=sweeps timeout 300 python scripts/time_tv.py hybrid upwind --shape 512 1 512 512 --rt 0 2>&1 | grep "tv " | tee -a $OUT/tv_times.txt
for lib in $EXTRA_LIBS; do
  PYTVB_LIB_PATH=$PWD/pytv-4d_b200/csrc/$lib timeout 300 python scripts/time_tv.py hybrid upwind central 2>&1 | grep "tv " | sed "s/^/$lib /" | tee -a $OUT/tv_times.txt
done
