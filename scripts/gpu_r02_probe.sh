#!/bin/bash
# Round-2 first GPU call: sanity of the round-1 tree, then the three probes the round-2 plan depends on
# (packed-fp32 issue rate, host-link bandwidth, the reference GPU path's own float32 error).
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.log
echo "== ffma2"; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2 scripts/ubench/ffma2.cu && timeout 120 /tmp/ffma2 2>&1 | tee $OUT/ffma2.txt
echo "== link"; timeout 300 python scripts/probe_link.py --out $OUT/link_n1.json 2>&1 | grep -v "^ \|^GPU\|^$" | head -20
echo "== ref fp32"; timeout 900 python scripts/probe_ref_fp32.py --out $OUT/ref_fp32.json 2>&1 | tail -40
echo "== tv timing"; timeout 300 python scripts/time_tv.py 2>&1 | grep "tv " | tee $OUT/tv_times.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee $OUT/smi.txt
