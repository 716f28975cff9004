#!/usr/bin/env python
"""Host-link micro-benchmark: pinned cudaMemcpyAsync H2D, D2H and both at once (what bounds bench.py's `e2e`).
Run alone (1 rank) or under torchrun (N concurrent ranks, one per GPU); rank 0 prints one JSON document.

    python scripts/probe_link.py [--mib 2048] [--out gpurun_out/link.json]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def measure(dev, nbytes, reps=5):
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(h2d, d2h):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize(dev)
        return time.perf_counter() - t0
    out = {}
    for name, a, b in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
        run(a, b)
        ts = [run(a, b) for _ in range(reps)]
        out[name + "_GBps_per_direction"] = nbytes / min(ts) / 1e9
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=2048)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo")
        dist.barrier()
    res = measure(dev, args.mib << 20)
    doc = {"ranks": world, "mib": args.mib, "rank0": res}
    if world > 1:
        allr = [None] * world
        dist.all_gather_object(allr, res)
        doc["per_rank"] = allr
        doc["aggregate_both_GBps_per_direction"] = sum(r["both_GBps_per_direction"] for r in allr)
    if rank == 0:
        try:
            import subprocess
            doc["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout[-3000:]
            doc["lscpu_numa"] = [l for l in subprocess.run(["lscpu"], capture_output=True, text=True).stdout.splitlines() if "NUMA" in l or "Model name" in l or l.startswith("CPU(s)")]
        except Exception as e:
            doc["topo_error"] = repr(e)
        s = json.dumps(doc, indent=1)
        print(s)
        if args.out:
            os.makedirs(os.path.dirname(args.out), exist_ok=True)
            open(args.out, "w").write(s)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
