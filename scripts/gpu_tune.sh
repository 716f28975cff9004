#!/bin/bash
# Time the CP bench with several builds of the library (tuning variants made by `make variant`).
# Usage: bash scripts/gpu_tune.sh TAG variant1 variant2 ...   ("default" = libpytv_b200.so)
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest (CP + slabs)"; timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -8 | tee $OUT/pytest_gpu.log
for v in "$@"; do
  if [ "$v" == "default" ]; then unset PYTVB_LIB_PATH; else export PYTVB_LIB_PATH=$PWD/pytv-4d_b200/csrc/libpytv_b200_$v.so; fi
  for gen in ${GENS:-2}; do
    PYTVB_GEN=$gen timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>$OUT/err_$v.log > $OUT/bench_${v}_g$gen.json
    python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_${v}_g$gen.json"))
    r=d["roofline"]
    print("%-10s gen%s  %.3f ms/step  dual %.3f ms (%.0f GB/s, %.3f)  primal %.3f ms (%.0f GB/s, %.3f)  iter frac %.3f  sm %s MHz" % ("$v","$gen",d["ms_per_step"],r["avg_launch_ms"],r["achieved"],r["frac"],r["pass_B"]["avg_launch_ms"],r["pass_B"]["achieved"],r["pass_B"]["frac"],r["iteration"]["frac"],d["clocks"]["sm_mhz"]))
except Exception as e:
    print("$v gen$gen FAILED", e); print(open("$OUT/err_$v.log").read()[-2000:])
PY
  done
done
