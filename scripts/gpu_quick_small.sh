#!/bin/bash
# Quick check after a change to the reductions / launch-bound paths: whole -m gpu suite, small-volume latencies, tv timings.
TAG=${1:-quicksmall}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -4 | tee $OUT/pytest_gpu.log
echo "== small volumes"; timeout 600 python scripts/bench_small.py --out $OUT/small.json > $OUT/small.log 2>&1; python -c "
import json; d=json.load(open('$OUT/small.json'))
for k,v in d.items(): print(k, {a:round(b,1) for a,b in v.items()} if isinstance(v,dict) else v)"
timeout 300 python scripts/time_tv.py hybrid upwind central 2>&1 | grep "tv " | tee $OUT/tv_times.txt
