#!/bin/bash
# Validate and time the single-launch iteration (generation 3) against the two-pass one.
TAG=${1:-fused}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -m gpu -q --tb=short -p no:cacheprovider -x -k "fused" 2>&1 | tail -6
for lag in ${LAGS:-1 2 3 4 6}; do
  PYTVB_FUSED=1 PYTVB_FUSED_LAG=$lag timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>$OUT/err_$lag.log > $OUT/bench_fused_lag$lag.json
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_fused_lag$lag.json")); print("fused lag=$lag  %.3f ms/step  value %.4g  energy %s" % (d["ms_per_step"], d["value"], d["config"]["energy_last"]))
except Exception as e:
    print("lag $lag FAILED", e); print(open("$OUT/err_$lag.log").read()[-1500:])
PY
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_twopass.json; python -c "
import json; d=json.load(open('$OUT/bench_twopass.json')); print('two-pass  %.3f ms/step  value %.4g  energy %s' % (d['ms_per_step'], d['value'], d['config']['energy_last']))"
