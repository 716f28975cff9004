#!/bin/bash
# Round 2: whole -m gpu suite with the auto tv path, small-volume latencies for both tv forms, bench C4 / C3 / C5 on one GPU.
TAG=${1:-r02h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -12 | tee $OUT/pytest_gpu.log
echo "== pytest -m gpu, tv forced through the tile kernel"; PYTVB_TV_PATH=tile timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -4 | tee $OUT/pytest_gpu_tile.log
echo "== small volumes"; timeout 600 python scripts/bench_small.py --out $OUT/small_auto.json > $OUT/small_auto.log 2>&1; tail -3 $OUT/small_auto.log
PYTVB_TV_PATH=sweeps timeout 600 python scripts/bench_small.py --out $OUT/small_sweeps.json > $OUT/small_sweeps.log 2>&1; tail -3 $OUT/small_sweeps.log
PYTVB_TV_PATH=tile timeout 600 python scripts/bench_small.py --out $OUT/small_tile.json > $OUT/small_tile.log 2>&1; tail -3 $OUT/small_tile.log
python -c "import bench; print(bench.lib_build_id())" > $OUT/lib_hash.txt
echo "== bench C4"; timeout 900 python bench.py 2> $OUT/bench.err | tee $OUT/bench_c4.json | cut -c1-300; tail -3 $OUT/bench.err
echo "== bench C3"; timeout 900 python bench.py --workload C3 --steps 200 2> $OUT/bench_c3.err | tee $OUT/bench_c3.json | cut -c1-300; tail -3 $OUT/bench_c3.err
echo "== bench C5"; timeout 900 python bench.py --workload C5 --steps 3 2> $OUT/bench_c5.err | tee $OUT/bench_c5.json | cut -c1-600; tail -3 $OUT/bench_c5.err
