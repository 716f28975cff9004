#!/bin/bash
# 2-GPU check of the peer-memory halo push: parity tests, then the bench with both halo transports.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multigpu.py -x -q > gpurun_out/p2p_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/p2p_tests.log
tail -5 gpurun_out/p2p_tests.log
for comm in nccl p2p nccl p2p; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --comm $comm --no-extras 2> gpurun_out/p2p_bench_$comm.err | tail -1 >> gpurun_out/p2p_bench_$comm.json
  echo "$comm rc=$?"; tail -c 600 gpurun_out/p2p_bench_$comm.err
done
python - <<'PY'
import json
for c in ("nccl","p2p"):
    for l in open("gpurun_out/p2p_bench_%s.json"%c):
        try: d=json.loads(l)
        except Exception: print(c,"bad line",l[:200]); continue
        print(c, d["ms_per_step"], d["value"], d.get("per_rank"), d["config"].get("halo_comm"))
PY
