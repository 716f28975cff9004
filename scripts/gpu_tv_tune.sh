#!/bin/bash
# TV sweeps: per-kernel times of tuning variants (C4 slab, hybrid); variant libraries are selected with PYTVB_LIB_PATH.
mkdir -p gpurun_out/tvtune
OUT=gpurun_out/tvtune
for tag in ${TAGS:-default}; do
  if [ $tag = default ]; then unset PYTVB_LIB_PATH; else export PYTVB_LIB_PATH=$PWD/pytv-4d_b200/csrc/libpytv_b200_$tag.so; fi
  timeout 300 python scripts/time_tv.py hybrid upwind 2>&1 | grep "tv " | sed "s/^/$tag: /" | tee -a $OUT/times.txt
  timeout 300 ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'tv_(norm|grad)_strip' -s 4 -c 2 --csv --log-file $OUT/ncu_$tag.csv python scripts/time_tv.py hybrid > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/ncu_$tag.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value")
for r in rows[1:]:
    print("$tag", r[ki][:24], r[mi], r[vi])
PY
done 2>&1 | tee $OUT/summary.txt
