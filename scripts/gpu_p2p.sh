#!/bin/bash
# 2-GPU check: multi-GPU parity tests (all halo transports), then the default bench (comm=auto) and the NCCL one.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multigpu.py -x -q > gpurun_out/p2p_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/p2p_tests.log
tail -5 gpurun_out/p2p_tests.log
for comm in auto nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --comm $comm --no-extras 2> gpurun_out/p2p_bench_$comm.err | tail -1 > gpurun_out/p2p_bench_$comm.json
  echo "$comm rc=$?"; grep -v "OMP_NUM\|\*\*\*\|^$\|NCCL version" gpurun_out/p2p_bench_$comm.err | tail -5
done
python - <<'PY'
import json
for c in ("auto","nccl"):
    for l in open("gpurun_out/p2p_bench_%s.json"%c):
        try: d=json.loads(l)
        except Exception: print(c,"bad line",l[:200]); continue
        print(c, d["ms_per_step"], d["value"], d.get("per_rank"), d["config"].get("halo_comm"), d["config"].get("energy_last"))
PY
