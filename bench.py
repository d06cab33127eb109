#!/usr/bin/env python
"""Headline benchmark: Beltrami grid-pt*steps/s (FP64), one `advance` time step per "step".

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W    # reference algorithm on the host cores (oracle port)

Workload (BASELINE.json configs[3], the configuration the metric is quoted on): Beltrami 512^3,
examples/beltrami_512.config (stepper cn2, Hou & Li filter, nnu = 3, prediss = 30, pretype vorch,
alpha = 0.1), synthetic analytic initial condition (beltrami.f90:162-181).  One JSON line on stdout.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "Beltrami grid-pt*steps/s (FP64)"
UNIT = "grid-pt*steps/s"
SWEEPS = {"cn2": 115, "impl-diff-rk4": 150}          # necessary 1-D transform sweeps per step


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([v.strip() for v in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax = float(r[2])
                for k, nme in enumerate(names):
                    if r[4 + k].lower().startswith("active"):
                        reasons.add(nme)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_run(n, steps, warmup, stepper="cn2"):
    """Reference algorithm on the host cores.  Preferred: oracle/ps3d_ref.cpp, the C++/OpenMP restatement with the
    reference's sweep structure (four transposes per 2-D FFT, reversed copies in diffx/diffy, stored N-sized tables,
    literal combine/decompose pairs in the stepper), all cores; fallback: the NumPy/SciPy oracle.
    Returns (grid-pt*steps/s, s/step, description)."""
    import math
    from ps3d_b200 import host
    lower = -0.5 * math.pi * np.ones(3)
    extent = math.pi * np.ones(3)
    try:
        from oracle.ps3d_ref import RefSolver
        r = RefSolver(n, n, n, lower, extent)
    except (OSError, FileNotFoundError, ValueError):
        r = None
    if r is not None:
        r.set_vorticity(host.beltrami_vorticity(n, n, n, lower, extent))
        for _ in range(warmup):
            r.advance(stepper=stepper)
        t0 = time.perf_counter()
        for _ in range(steps):
            r.advance(stepper=stepper)
        dt = time.perf_counter() - t0
        threads = r.threads
        r.close()
        return n ** 3 * steps / dt, dt / steps, ("C++/OpenMP restatement of the reference algorithm (oracle/ps3d_ref.cpp: 4 transposes "
                                                  f"per 2-D FFT, literal combine/decompose pairs; not the Fortran build), {threads} OpenMP threads")
    from oracle import ps3d_oracle as O
    s = O.beltrami_setup(n)
    t = 0.0
    for _ in range(warmup):
        t, _ = s.advance(t, 100.0, stepper, literal=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        t, _ = s.advance(t, 100.0, stepper, literal=True)
    dt = time.perf_counter() - t0
    return n ** 3 * steps / dt, dt / steps, "NumPy/SciPy port of the reference algorithm (literal steppers), scipy.fft on all cores"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.ref_n
    val, sec, how = cpu_reference_run(n, args.steps, args.warmup, args.stepper)
    cores = os.cpu_count()
    sample = f"Beltrami {n}^3 {args.stepper} steps (bounded sample of the {args.n}^3 workload), {how}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"Beltrami {args.n}^3 {args.stepper} (examples/beltrami_512.config)", "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--grid", "--n", dest="n", type=int, default=512, help="grid size (nx = ny = nz)")
    ap.add_argument("--nz", type=int, default=0, help="vertical cells if different from --grid (dev runs)")
    ap.add_argument("--stepper", default="cn2")
    ap.add_argument("--ref-n", type=int, default=128, help="grid of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path)")
    os.environ.setdefault("CUDA_DEVICE_ORDER", "PCI_BUS_ID")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import ps3d_b200
    from ps3d_b200 import host
    lib = ps3d_b200.load()
    # NOTE: ps3d_cuda_init selects device rank % ndev; with one process per GPU we present exactly
    # this process's device as device 0 through the rank argument below.
    n = args.n
    lower = -0.5 * math.pi * np.ones(3)
    extent = math.pi * np.ones(3)
    os.environ["PS3D_DEVICE"] = str(local_rank)
    nccl_id = None
    if world > 1:
        # slab decomposition over the GPUs of the box: x-slabs in physical space, ky-slabs in spectral space,
        # one NCCL all-to-all per 2-D FFT inside the library (its own communicator, id made here)
        box = [torch.cuda.nccl.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        nccl_id = box[0]
    nzz = args.nz or n
    solver = host.Solver(lib, n, n, nzz, lower, extent, stepper=args.stepper, rank=rank, nranks=world, nccl_id=nccl_id)
    nxl = n // world
    vor_np = host.beltrami_vorticity(n, n, nzz, lower, extent, x0=rank * nxl, x1=(rank + 1) * nxl)
    vor_pinned = torch.from_numpy(vor_np).pin_memory()
    vor_host = vor_pinned.numpy()
    solver.setup_fields(vor_host)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        solver.advance()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    l0 = lib.kernel_launches()
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        solver.advance()
        dev_ms += lib.last_advance_ms()          # CUDA events on the library's stream around the whole step
    barrier()
    wall = time.perf_counter() - t0
    launches = lib.kernel_launches() - l0
    clocks = sampler.stop()
    ms_step = dev_ms / args.steps
    if world > 1:
        tt = torch.tensor([ms_step, wall], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_step, wall = float(tt[0]), float(tt[1])

    # ---- end-to-end through the C ABI with host buffers (H2D of the step's input, D2H of its result) ----
    h2d = vor_host.nbytes
    e2e_steps = max(2, min(args.steps, 5))
    solver.lib.upload_vorticity(vor_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        lib.upload_vorticity(vor_host)           # 3 fields, pinned host memory -> HBM, decomposed on device
        solver.t = 0.0
        solver.advance()                         # diag_out[16] comes back to the host
        d = lib.diagnostics()                    # KE / enstrophy / helicity read back
    barrier()
    e2e_sec = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        tt = torch.tensor([e2e_sec], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_sec = float(tt[0])

    # ---- per-kernel device times (CUDA events on the library's stream), roofline of the dominant one ----
    peak, peak_src = peaks()
    N = n * n * (nzz + 1) // world                         # array elements per field on this rank
    kinfo = [("line_fwd_y", 16), ("line_fwd_x", 16), ("line_inv_x", 16), ("line_inv_y", 16),
             ("vor2vel_columns", 8 * 16), ("source_columns", 5 * 16)]
    kernels = {}
    for w, (name, bpp) in enumerate(kinfo):
        lib.time_kernel(w, 2)
        ms = lib.time_kernel(w, 10)
        kernels[name] = {"ms": ms, "alg_bytes": bpp * N * 8 // 8, "GBs": bpp * N / (ms * 1e-3) / 1e9}
    # share of the step: launches per cn2 step = 3 vor2vel + 3 source column kernels, 3*18 + 10 line sweeps
    mult = 3 if args.stepper == "cn2" else 4
    share = {"vor2vel_columns": mult * kernels["vor2vel_columns"]["ms"],
             "source_columns": mult * kernels["source_columns"]["ms"],
             "line_sweeps": (mult * 18 + 10) * np.mean([kernels[k]["ms"] for k in list(kernels)[:4]])}
    dom = max(share, key=share.get)
    if dom == "line_sweeps":
        dom = "line_fwd_y"
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture of this command (per launch),
    # only quoted for the configuration it was captured on
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01j_ncu_kernels.json")
    if os.path.exists(tpath) and n == 512 and world == 1:
        t = json.load(open(tpath)).get(dom if dom in ("vor2vel_columns", "source_columns") else "line_fwd_y")
        if t:
            traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
    roof = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["GBs"], "peak": peak, "unit": "GB/s",
            "frac": kernels[dom]["GBs"] / peak, "traffic": traffic, "peak_source": peak_src,
            "alg_bytes_per_launch": kernels[dom]["alg_bytes"], "ms_per_launch": kernels[dom]["ms"]}
    step_bytes = SWEEPS[args.stepper] * 16 * n * n * (nzz + 1)   # whole job, peak = P x one GPU
    value = n * n * nzz / (ms_step * 1e-3)                      # whole job: the grid is split over the ranks
    n_a2a, sent = lib.comm_stats()
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"Beltrami {n}^3 {args.stepper} (examples/beltrami_512.config), analytic IC k=l=2 m=1",
                   "grid": [n, n, n], "stepper": args.stepper, "filtering": "Hou & Li", "nnu": 3, "prediss": 30.0,
                   "parallelism": "1 GPU" if world == 1 else
                   f"slab{world}: x-slabs / ky-slabs, one NCCL all-to-all per 2-D FFT",
                   "l2": "inputs larger than L2 (each field %.2f GB per GPU vs 126 MB L2)" % (N * 8 / 1e9)},
        "roofline": roof,
        "step_roofline": {"alg_bytes_per_step": step_bytes, "achieved": step_bytes / (ms_step * 1e-3) / 1e9,
                          "peak": peak * world, "unit": "GB/s", "frac": step_bytes / (ms_step * 1e-3) / 1e9 / (peak * world)},
        "kernels": kernels, "step_share_ms": share,
        "e2e": {"value": n * n * nzz / e2e_sec, "unit": UNIT, "h2d_bytes_per_step": int(h2d) * world,
                "d2h_bytes_per_step": (16 + 8) * 8, "ms_per_step": e2e_sec * 1e3},
        "comm": {"alltoalls_total": int(n_a2a), "bytes_sent_per_rank_total": sent},
        "gpu_launches": int(launches), "wall_ms_per_step": wall / args.steps * 1e3, "clocks": clocks,
        "diag": {k: float(v) for k, v in d.items()},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count()
        val, sec, how = cpu_reference_run(args.ref_n, 3, 1, args.stepper)
        out["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"Beltrami {args.ref_n}^3 {args.stepper}, 3 steps after 1 warm-up, {how}"}
    solver.close()
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
