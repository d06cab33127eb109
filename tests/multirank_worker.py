"""Worker of the world_size-2 gloo test: one rank of the slab decomposition on the CPU block emulator,
with the all-to-all / all-reduce supplied by torch.distributed (gloo) through ps3d_cuda_set_transport."""
import ctypes as C
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ps3d_oracle as O                     # noqa: E402
from ps3d_b200.lib import PS3DLib, ALLTOALL_FN, ALLREDUCE_FN   # noqa: E402


def rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def main():
    emu = sys.argv[1]
    stepper = sys.argv[2]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    buoyancy = len(sys.argv) > 4 and sys.argv[4] == "buoyancy"      # ENABLE_BUOYANCY paths on the slab decomposition
    nx, ny, nz = n, n, n
    lower = np.array([-0.5 * math.pi] * 3)
    extent = np.array([math.pi] * 3)
    use_cuda = (emu == "cuda")
    if use_cuda:
        # product library, NCCL transport inside the library (its own communicator)
        import ps3d_b200
        os.environ["PS3D_DEVICE"] = os.environ.get("LOCAL_RANK", "0")
        lib = ps3d_b200.load()
        box = [torch.cuda.nccl.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        lib.init(nx, ny, nz, lower, extent, rank, world, box[0])
    else:
        lib = PS3DLib(emu)
        lib.init(nx, ny, nz, lower, extent, rank, world)

    def alltoall(send, recv, nbytes, user):
        n = nbytes // 8
        s = torch.from_numpy(np.ctypeslib.as_array((C.c_double * (n * world)).from_address(send)))
        r = torch.from_numpy(np.ctypeslib.as_array((C.c_double * (n * world)).from_address(recv)))
        outs = list(r.split(n))
        ins = [t.contiguous() for t in s.split(n)]
        dist.all_to_all(outs, ins) if dist.get_backend() != "gloo" else _a2a_gloo(outs, ins)
        return 0

    def _a2a_gloo(outs, ins):
        reqs = []
        for d in range(world):
            if d == rank:
                outs[d].copy_(ins[d])
            else:
                reqs.append(dist.isend(ins[d], d))
                reqs.append(dist.irecv(outs[d], d))
        for q in reqs:
            q.wait()

    def allreduce(buf, n, op, user):
        t = torch.from_numpy(np.ctypeslib.as_array(buf, shape=(n,)))
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op else dist.ReduceOp.SUM)
        return 0

    if not use_cuda:
        lib.set_transport(ALLTOALL_FN(alltoall), ALLREDUCE_FN(allreduce))
    lib.init_inversion("Hou & Li")
    ref = O.PS3D(nx, ny, nz, lower, extent)
    vor = np.random.default_rng(5).uniform(-1, 1, (3, nx, ny, nz + 1))
    ref.set_vorticity(vor)
    nxl, nyl = nx // world, ny // world
    xs = slice(rank * nxl, (rank + 1) * nxl)
    kys = PS3DLib.paired_ky(ny)[rank * nyl:(rank + 1) * nyl]
    if buoyancy:
        f_cor, bfsq = (0.0, 0.1, 0.2), 0.7
        buoy = np.random.default_rng(6).uniform(-0.5, 0.5, (nx, ny, nz + 1))
        ref.f_cor = np.asarray(f_cor)
        ref.enable_buoyancy(buoy, bfsq)
        lib.set_physics(f_cor, bfsq)
        lib.enable_buoyancy()
    lib.upload_vorticity(np.ascontiguousarray(vor[:, xs]))
    if buoyancy:
        lib.upload_buoyancy(np.ascontiguousarray(buoy[xs]))
    lib.vor2vel()
    errs = {}
    for name in ("vor", "vel"):
        errs[name] = rel(lib.download3(name), getattr(ref, name)[:, xs])
    for name in ("svor", "svel"):
        errs[name] = rel(lib.download3(name), getattr(ref, name)[:, :, kys])
    d = lib.diagnostics()
    errs["ke"] = abs(d["ke"] - ref.get_kinetic_energy()) / ref.get_kinetic_energy()
    errs["en"] = abs(d["en"] - ref.get_enstrophy()) / ref.get_enstrophy()
    lib.init_diffusion(d["ke"], d["en"])
    if buoyancy:
        lib.init_diffusion_buoyancy(d["ke"], d["en"], 3, 20.0, "Kolmogorov", "roll-mean-bfmax", 2)
        ref.init_diffusion_buoyancy(ref.get_kinetic_energy(), ref.get_enstrophy(), 3, 20.0)
    lib.stepper_setup(stepper)
    t = tr = 0.0
    for _ in range(2):
        t, dt, diag = lib.advance(t, 100.0)
        tr, dtr = ref.advance(tr, 100.0, stepper, literal=True, bpretype="roll-mean-bfmax", bwin=2)
        errs["dt"] = max(errs.get("dt", 0.0), abs(dt - dtr) / dtr)
        if buoyancy:
            errs["bfmax"] = max(errs.get("bfmax", 0.0), abs(diag["bfmax"] - ref.diag["bfmax"]) / ref.diag["bfmax"])
            errs["sbuoy"] = max(errs.get("sbuoy", 0.0), rel(lib.download("sbuoy"), ref.sbuoy[:, kys]))
        for k in ("vortmax", "vortrms", "vorch", "ggmax", "umax", "usggmax", "lsggmax"):
            errs[k] = max(errs.get(k, 0.0), abs(diag[k] - ref.diag[k]) / abs(ref.diag[k]))
    errs["svor_after"] = rel(lib.download3("svor"), ref.svor[:, :, kys])
    # the 40 field statistics (reduced over the ranks inside the library) for the state after the last step
    lib.vor2vel(); lib.adapt(t, 100.0)
    ref.vor2vel(); ref.adapt(tr, 100.0)
    got, want = lib.field_stats(), ref.field_stats()
    for k, w in want.items():
        if k in ("romin", "romax"):
            continue
        errs["stat_" + k] = abs(got[k] - w) / max(abs(w), 1e-3)
    n_a2a, sent = lib.comm_stats()
    lib.finalise()
    worst = max(errs.values())
    print(f"rank {rank} worst {worst:.3e} alltoalls {n_a2a} sent {sent:.0f} B", {k: f"{v:.1e}" for k, v in errs.items()}, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if worst < 1e-12 else 1)


if __name__ == "__main__":
    main()
