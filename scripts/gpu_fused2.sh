#!/bin/bash
TAG=${1:-fused2}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -m gpu -q --tb=short -p no:cacheprovider -x -k "fused" 2>&1 | tail -4
for v in "$@"; do
  if [ "$v" == "default" ]; then unset PYTVB_LIB_PATH; else export PYTVB_LIB_PATH=$PWD/pytv-4d_b200/csrc/libpytv_b200_$v.so; fi
  for lag in ${LAGS:-1 2 3 4}; do
    PYTVB_FUSED=1 PYTVB_FUSED_LAG=$lag timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>$OUT/err.log > $OUT/b.json
    python -c "
import json
try:
    d=json.load(open('$OUT/b.json')); print('$v fused lag=$lag  %.3f ms/step  energy %s' % (d['ms_per_step'], d['config']['energy_last']))
except Exception as e: print('$v lag $lag FAILED', e); print(open('$OUT/err.log').read()[-800:])"
  done
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/b2.json; python -c "
import json; d=json.load(open('$OUT/b2.json')); print('$v two-pass  %.3f ms/step' % d['ms_per_step'])"
done
