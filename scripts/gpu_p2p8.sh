#!/bin/bash
# N-GPU bench with both halo transports (same seeds: energy_last must agree bitwise between the two).
N=${1:-8}
mkdir -p gpurun_out
du -sh . 2>/dev/null | tail -1
for comm in nccl p2p; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 --comm $comm --no-extras 2> gpurun_out/p2p${N}_bench_$comm.err | tail -1 > gpurun_out/p2p${N}_bench_$comm.json
  echo "$comm rc=$?"; grep -v "OMP_NUM\|\*\*\*\|^$\|NCCL version" gpurun_out/p2p${N}_bench_$comm.err | tail -5
done
python - <<PY
import json
for c in ("nccl","p2p"):
    for l in open("gpurun_out/p2p${N}_bench_%s.json"%c):
        try: d=json.loads(l)
        except Exception: print(c,"bad line",l[:200]); continue
        pr=d.get("per_rank") or {}
        print(c, "ms/step %.3f"%d["ms_per_step"], "value %.4g"%d["value"], "energy", repr(d["config"].get("energy_last")), d["config"].get("halo_comm"))
        for k,v in pr.items(): print("   ",k,["%.2f"%x for x in v])
PY
