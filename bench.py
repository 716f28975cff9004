#!/usr/bin/env python
"""Benchmark of the TV hot path: Chambolle-Pock iteration voxel-updates/s (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (sm_100a kernels)
    python bench.py --impl reference [--gpus N] [--steps K] ...     the reference's CPU path on the host cores

Workload (BASELINE.json configs[3], the configuration the metric is quoted on): per-GPU z-slab
(128, 4, 1024, 1024) float32 of a (128*N, 4, 1024, 1024) dynamic-CT volume, hybrid scheme with time
regularisation reg_time = 2^-5 (Nd = 8 difference components), ROF-form Chambolle-Pock iteration = dual
pass + primal pass, energy tracked every iteration.  Weak scaling: Nz grows with N, slabs exchange one-plane
halos before each pass (NCCL send/recv) and all-reduce two doubles per iteration for the energy.

A "step" is one iteration over the whole volume.  `value` = voxels * K / t with the state resident in HBM
(t = CUDA-event time of K iterations, max over ranks).  `e2e` = the same iteration driven through the public
host-buffer call `CPSolver.step_host_async`: every step uploads the data term x0 from pinned host memory, runs
the iteration, downloads the current image x and the energy (steps pipelined over two copy streams).  The state per GPU (23.6 GB) is far larger than the
126 MB L2, so no L2 flush is needed between iterations.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cp_iter_voxel_updates_per_s"
UNIT = "voxel-updates/s"
SLAB = (128, 4, 1024, 1024)          # per-GPU slab of BASELINE config 4
REG_TIME = 2.0 ** -5
LAM = 0.1
WORKLOAD = "C4 dynamic-CT slab (Nz=128*N_gpus, M=4, N=1024) f32, hybrid scheme, reg_time=2^-5 (Nd=8), CP-ROF iteration"


def config_dict(n_gpus, extra=None):
    cfg = {"workload": WORKLOAD, "slab_per_gpu": list(SLAB), "scheme": "hybrid", "Nd": 8, "reg_time": REG_TIME, "lam": LAM,
           "variant": "rof", "parallelism": "z-slabs x%d, 1-plane halos" % n_gpus,
           "l2_policy": "working set 23.6 GB per GPU >> 126 MB L2, no flush needed"}
    if extra:
        cfg.update(extra)
    return cfg


# --------------------------------------------------------------------------------------------------------
# CPU arms (oracle / reference): test-infrastructure code, used only as the reported baseline
def _cpu_ops():
    """(D, D_T, l21, kind): the unmodified reference numpy functions when a copy travels with the repo
    (baseline/_ref, git-ignored), else the oracle port."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(os.path.join(ref_dir, "pytv")) and os.environ.get("PYTVB_BENCH_FORCE_PORT") != "1":
        try:
            import importlib.util
            import warnings
            warnings.filterwarnings("ignore", category=SyntaxWarning)
            spec = importlib.util.spec_from_file_location("_ref_tv_operators_CPU", os.path.join(ref_dir, "pytv", "tv_operators_CPU.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod.D_hybrid, mod.D_T_hybrid, mod.compute_L21_norm, "reference"
        except Exception:
            pass
    from oracle import tv_oracle as orc
    return (lambda x, **kw: orc.D(x, "hybrid", **kw)), (lambda p, **kw: orc.D_T(p, "hybrid", **kw)), orc.l21, "port"


def _cpu_cp_iteration(D, DT, l21, state, x0, lam, sigma, tau, kw):
    """CP-ROF iteration from the CPU operators (the loop of README.md:145-157 in its over-relaxed form)."""
    x, xbar, y = state
    Dxb = D(xbar, **kw)
    pa = y + sigma * Dxb
    y = pa / np.maximum(1.0, np.sqrt(np.sum(pa ** 2, axis=1, keepdims=True)) / lam)
    x_new = (x - tau * DT(y, **kw) + tau * x0) / (1.0 + tau)
    xbar = x_new + (x_new - x)
    energy = 0.5 * np.sum(np.square(x_new - x0)) + lam * l21(Dxb)
    return (x_new, xbar, y), float(energy)


def _cpu_worker(args):
    """One process of the CPU arm: `steps` timed iterations on its own sample slab."""
    seed, shape, steps, warmup = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    D, DT, l21, kind = _cpu_ops()
    rs = np.random.RandomState(seed)
    x0 = (rs.rand(*shape) + 0.05 * rs.randn(*shape)).astype(np.float32)
    kw = dict(reg_time=REG_TIME)
    y = np.zeros((shape[0], 8) + shape[1:], np.float32)
    state = (x0.copy(), x0.copy(), y)
    sigma, tau = np.float32(0.5), np.float32(1.0 / (4.0 * (3.0 + REG_TIME) + 1.0))
    for _ in range(warmup):
        state, e = _cpu_cp_iteration(D, DT, l21, state, x0, np.float32(LAM), sigma, tau, kw)
    t0 = time.perf_counter()
    for _ in range(steps):
        state, e = _cpu_cp_iteration(D, DT, l21, state, x0, np.float32(LAM), sigma, tau, kw)
    return time.perf_counter() - t0, kind, e


def _cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_baseline_single(sample_shape=(8, 4, 512, 512), steps=7, warmup=1):
    """Single-process run of the CPU implementation, as shipped (numpy element-wise kernels: one busy core)."""
    dt, kind, _ = _cpu_worker((1000, sample_shape, steps, warmup))
    vox = int(np.prod(sample_shape))
    return {"value": vox * steps / dt, "unit": UNIT, "cores": 1, "kind": kind, "host_cpu": _cpu_model(), "host_cpu_count": os.cpu_count(),
            "sample": "%d CP-ROF iterations on a %s float32 sample slab of the same workload (hybrid, reg_time=2^-5, Nd=8), %.1f s"
                      % (steps, "x".join(map(str, sample_shape)), dt)}


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path on all host cores of the box: one
    process per core, each iterating on its own sample slab (independent z-slabs, which is how the CPU path
    would be sharded); voxel-updates/s is the aggregate."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    sample = (4, 4, 512, 512)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    # bound the run to a few minutes: one iteration of this sample takes ~1.5 s on one core
    steps = min(steps, 20)
    warmup = min(warmup, 3)
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(procs) as pool:
        res = pool.map(_cpu_worker, [(1000 + p, sample, steps, warmup) for p in range(procs)])
    wall = time.perf_counter() - t0
    t_max = max(r[0] for r in res)
    kind = res[0][1]
    vox = int(np.prod(sample)) * procs
    value = vox * steps / t_max
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": 1e3 * t_max / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": config_dict(args.gpus, {"note": "CPU arm: bounded sample of the workload, %d processes x slab %s" % (procs, list(sample))}),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind, "host_cpu": _cpu_model(), "host_cpu_count": os.cpu_count(),
                             "sample": "%d processes (of %d host cores), each %d CP-ROF iterations on its own %s float32 slab; wall %.1f s"
                                       % (procs, cores, steps, "x".join(map(str, sample)), wall)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampling of SM clocks and throttle reasons during the timed region (B200_PROFILING.md)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        try:
            rows = [r.split(",") for r in open(self.path).read().strip().splitlines() if r.strip()]
            sm = sorted(float(r[1]) for r in rows)
            reasons = set()
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for r in rows:
                for name, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(rows[0][2]) if rows else None,
                   "reasons": sorted(reasons), "samples": len(rows), "power_w_max": max(float(r[3]) for r in rows) if rows else None}
            os.unlink(self.path)
        except Exception:
            pass
        return out


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def recorded_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture, if any."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    import pytv_b200 as pytv
    from pytv_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout: keep stdout clean for the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    n_gpus = world
    if args.gpus != world and rank == 0:
        print("bench.py: --gpus %d but WORLD_SIZE=%d; running on %d" % (args.gpus, world, world), file=sys.stderr)
    K, W = args.steps, max(3, args.warmup)
    shape = tuple(args.slab) if args.slab else SLAB
    V_local = int(np.prod(shape))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- synthetic data, seeded per rank (uniform [0,1) + 0.05 randn, seed 1000+rank)
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    x0 = torch.rand(shape, generator=g, device=dev) + 0.05 * torch.randn(shape, generator=g, device=dev)
    comm = args.comm if world > 1 else "nccl"
    solver = pytv.CPSolver(x0, lam=LAM, scheme="hybrid", variant="rof", reg_time=REG_TIME, distributed=(world > 1),
                           z_offset=rank * shape[0], Nz_global=world * shape[0], comm=comm)
    del x0
    assert solver.Nd == 8 or args.slab
    lib = _lib.lib()

    # ---- device-resident timing
    for _ in range(W):
        solver.step()
        solver.energy()
        if world > 1:
            solver.energy_result(solver.energy_async())
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    time.sleep(0.3)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = lib.pytvb_launch_count()
    barrier()
    start.record()
    for k in range(K):
        # the solver's step(), spelled out so that the two passes can be bracketed by events on their stream
        # (sharded: a pass = boundary planes, halo send/recv started, interior planes)
        ev[k][0].record()
        if solver.fused:
            solver.iterations -= 1
            solver.step(1)                 # one launch: pass B tiles lag pass A tiles inside the same kernel
            ev[k][1].record()
        else:
            solver._pass_A()
            ev[k][1].record()
            solver._pass_B()
        ev[k][2].record()
        solver.iterations += 1
        if world > 1:
            pending_energy = solver.energy_async()     # scalar all-reduce for the energy, every iteration (own communicator)
    if world > 1:
        solver.energy_result(pending_energy)           # the last all-reduce lands inside the timed region
    end.record()
    barrier()
    launches = lib.pytvb_launch_count() - launches0
    t_ms = start.elapsed_time(end)
    clocks = sampler.stop() if sampler else None
    energy = solver.energy()
    dual_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / K
    primal_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / K
    per_rank = None
    if world > 1:
        mine = torch.tensor([t_ms, dual_ms, primal_ms], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"total_ms": [float(a[0]) for a in allr], "pass_A_ms": [float(a[1]) for a in allr], "pass_B_ms": [float(a[2]) for a in allr]}
        t = mine.clone()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_ms, dual_ms, primal_ms = t.tolist()
    value = V_local * world * K / (t_ms * 1e-3)

    # ---- end to end through the host-buffer API (pinned host memory in and out, every step)
    # step_host_async pipelines the three legs of a step (upload of x0 | the two passes | download of x and the
    # energy) across consecutive steps; every step still uploads its data and downloads its result inside the
    # timed region, and the region ends only when the last download has landed.
    x0_host = torch.empty(shape, dtype=torch.float32).pin_memory()
    x0_host.copy_(solver.x0)
    x_hosts = [torch.empty(shape, dtype=torch.float32).pin_memory() for _ in range(2)]
    for _ in range(2):
        solver.wait(solver.step_host_async(x0_host, x_hosts[0]))
    barrier()
    t0 = time.perf_counter()
    prev = None
    for k in range(K):
        ticket = solver.step_host_async(x0_host, x_hosts[k % 2])
        if prev is not None:
            solver.wait(prev)
        prev = ticket
    e2e_energy = solver.wait(prev)
    barrier()
    e2e_s = time.perf_counter() - t0
    # the unpipelined call, for reference
    t1 = time.perf_counter()
    for _ in range(3):
        solver.step_host(x0_host, x_hosts[0])
    torch.cuda.synchronize()
    e2e_sync_ms = (time.perf_counter() - t1) / 3 * 1e3
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    e2e_value = V_local * world * K / e2e_s
    img_bytes = V_local * 4

    # ---- informational: the opt-in reduced-precision mode (dual field stored as normalised half; NOT the parity path)
    extras = {}
    if world == 1 and not args.no_extras:
        hs = pytv.CPSolver(solver.x0, lam=LAM, scheme="hybrid", variant="rof", reg_time=REG_TIME, dual_dtype=torch.float16)
        for _ in range(W):
            hs.step()
        torch.cuda.synchronize()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        hs.step(K)
        h1.record()
        torch.cuda.synchronize()
        hms = h0.elapsed_time(h1) / K
        hbytes = (6.0 * hs.Nd + 20.0) * V_local
        extras["half_precision_dual_storage"] = {"ms_per_step": hms, "value": V_local / (hms * 1e-3), "unit": UNIT, "bytes_per_voxel": 6 * hs.Nd + 20,
                                                 "achieved_GBps": hbytes / (hms * 1e-3) / 1e9,
                                                 "note": "CPSolver(dual_dtype=torch.float16): opt-in, max error ~3e-4 on [0,1] data; not the headline"}
        del hs

    if rank == 0:
        Nd = solver.Nd
        peak, peak_src = measured_hbm_peak()
        dual_bytes = 4.0 * (2 * Nd + 1) * V_local       # read xbar, read y, write y  (pass A of 4(3Nd+5))
        primal_bytes = 4.0 * (Nd + 4) * V_local         # read y, x, x0; write x, xbar
        achieved = dual_bytes / (dual_ms * 1e-3) / 1e9
        traffic = recorded_traffic()
        if solver.fused:
            # single-launch iteration: y is read from DRAM once (pass A) and found in L2 by pass B
            fused_bytes = 4.0 * (2 * Nd + 5) * V_local
            it_ms = t_ms / K
            ach = fused_bytes / (dual_ms * 1e-3) / 1e9
            roofline = {"bound": "hbm", "kernel": "cp_fused_kernel (pass A + lagging pass B in one launch)", "achieved": ach, "peak": peak, "unit": "GB/s",
                        "frac": ach / peak, "peak_source": peak_src, "algorithmic_bytes_per_launch": fused_bytes, "avg_launch_ms": dual_ms,
                        "traffic": traffic.get("cp_fused_kernel") if traffic else None,
                        "note": "algorithmic bytes of this kernel are 4(2Nd+5) per voxel (x-bar, y read; y written; x, x0 read; x, x-bar written); "
                                "the two-pass formulation of SURVEY 8d moves 4(3Nd+5)",
                        "two_pass_equivalent": {"algorithmic_bytes": dual_bytes + primal_bytes,
                                                "achieved": (dual_bytes + primal_bytes) / (it_ms * 1e-3) / 1e9,
                                                "frac": (dual_bytes + primal_bytes) / (it_ms * 1e-3) / 1e9 / peak}}
        else:
          roofline = {"bound": "hbm", "kernel": "cp_dual_kernel (pass A)", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": dual_bytes, "avg_launch_ms": dual_ms,
                    "traffic": traffic.get("cp_dual_kernel") if traffic else None,
                    "pass_B": {"kernel": "cp_primal_kernel", "achieved": primal_bytes / (primal_ms * 1e-3) / 1e9,
                               "frac": primal_bytes / (primal_ms * 1e-3) / 1e9 / peak, "avg_launch_ms": primal_ms,
                               "traffic": traffic.get("cp_primal_kernel") if traffic else None},
                    "iteration": {"algorithmic_bytes": dual_bytes + primal_bytes, "achieved": (dual_bytes + primal_bytes) * K / (t_ms * 1e-3) / 1e9,
                                  "frac": (dual_bytes + primal_bytes) * K / (t_ms * 1e-3) / 1e9 / peak}}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": K, "warmup": W, "ms_per_step": t_ms / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_dict(n_gpus, {"slab_per_gpu": list(shape), "energy_last": energy,
                                               "halo_comm": "none (one GPU)" if world == 1 else
                                               ("p2p (kernels store boundary planes into the neighbours' halo buffers over NVLink)"
                                                if solver._peer is not None else "nccl send/recv")}),
                "roofline": roofline,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": img_bytes * world, "d2h_bytes_per_step": (img_bytes + 16) * world,
                        "ms_per_step": 1e3 * e2e_s / K, "ms_per_step_unpipelined": e2e_sync_ms,
                        "what": "CPSolver.step_host_async/wait: every step uploads x0 from pinned host memory, runs one iteration, "
                                "downloads x and the energy; consecutive steps are pipelined over two copy streams"},
                "gpu_launches": int(launches), "clocks": clocks}
        if extras:
            line["extras"] = extras
        if per_rank:
            line["per_rank"] = per_rank      # event times of every rank (the headline uses the max)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_single()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--comm", choices=["auto", "nccl", "p2p"], default=os.environ.get("PYTVB_BENCH_COMM", "auto"),
                    help="N > 1: halo planes by NCCL send/recv between the passes, or pushed by the kernels into peer memory")
    ap.add_argument("--slab", type=int, nargs=4, default=None, help="override the per-GPU slab (Nz M Ni Nj); debugging only")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU baseline leg (profiling runs)")
    ap.add_argument("--no-extras", action="store_true", help="skip the informational reduced-precision measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
