#!/bin/bash
# compute-sanitizer over the small-shape GPU tests: tile kernel (both forms), the single-GPU halo-push schedule.
TAG=${1:-san}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for form in 2 1; do
for tool in racecheck synccheck memcheck; do
  PYTVB_TILE_FORM=$form timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_parity.py tests/test_gpu_push_single.py -m gpu -q -x -p no:cacheprovider -k "readme_volume or delta or push or small_goldens_float32" 2>&1 | grep -v "^$" > $OUT/sanitizer_${tool}_form$form.full
  grep -c "Race reported\|Error:" $OUT/sanitizer_${tool}_form$form.full
  (grep -A3 "Race reported\|Error:" $OUT/sanitizer_${tool}_form$form.full | head -40; tail -3 $OUT/sanitizer_${tool}_form$form.full) > $OUT/sanitizer_${tool}_form$form.log
  tail -2 $OUT/sanitizer_${tool}_form$form.log
  rm -f $OUT/sanitizer_${tool}_form$form.full
done
done
