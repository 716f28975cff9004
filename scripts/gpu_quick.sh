#!/bin/bash
# Quick round-end confirmation on one GPU: smoke, the whole -m gpu suite, the default bench line.
OUT=gpurun_out/${1:-quick}
mkdir -p $OUT
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $OUT/smoke.log
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -2 | tee $OUT/pytest_gpu.log
timeout 600 python bench.py --no-extras 2> $OUT/bench.err | tee $OUT/bench.json | cut -c1-200
timeout 120 python scripts/time_tv.py 2>&1 | grep "tv " | tee $OUT/tv_times.txt
