"""Host-side mirror of the reference's run loop over the C ABI.

`Solver` plays the part of `ps3d.f90` (pre_run / run) for tests and benchmarks:
it follows utils.f90:136-184 (`setup_fields`) and ps3d.f90:107-126 (`run`)
call for call, with every numerical procedure replaced by its
`ps3d_cuda_*` entry point.  Names follow the reference (stepper, filtering,
vor_visc%..., time%...).
"""
from __future__ import annotations

import math

import numpy as np


def beltrami_vorticity(nx, ny, nz, lower, extent, k=2, l=2, m=1, x0=0, x1=None):
    """Initial condition of examples/beltrami*.nml (beltrami.f90:141-181); x0:x1 selects an x-slab."""
    lower = np.asarray(lower, dtype=np.float64)
    extent = np.asarray(extent, dtype=np.float64)
    dx = extent / np.array([nx, ny, nz], dtype=np.float64)
    kk, ll, mm = float(k), float(l), float(m)
    alpha = math.sqrt(kk ** 2 + ll ** 2 + mm ** 2)
    fk2l2 = alpha / float(k ** 2 + l ** 2)
    x1 = nx if x1 is None else x1
    x = (lower[0] + dx[0] * np.arange(x0, x1))[:, None, None]
    y = (lower[1] + dx[1] * np.arange(ny))[None, :, None]
    z = (lower[2] + dx[2] * np.arange(nz + 1))[None, None, :]
    cosmz, sinmz = np.cos(mm * z), np.sin(mm * z)
    s, c = np.sin(kk * x + ll * y), np.cos(kk * x + ll * y)
    vor = np.empty((3, x1 - x0, ny, nz + 1))
    vor[0] = fk2l2 * (kk * mm * sinmz - ll * alpha * cosmz) * s
    vor[1] = fk2l2 * (ll * mm * sinmz + kk * alpha * cosmz) * s
    vor[2] = alpha * cosmz * c
    return vor


class Solver:
    def __init__(self, lib, nx, ny, nz, lower, extent, *, stepper="cn2", filtering="Hou & Li", nnu=3, prediss=30.0,
                 pretype="vorch", length_scale="Kolmogorov", roll_mean_win_size=1000, alpha=0.1, limit=100.0,
                 rank=0, nranks=1, nccl_id=None):
        self.lib = lib
        self.opts = dict(stepper=stepper, filtering=filtering, nnu=nnu, prediss=prediss, pretype=pretype,
                         length_scale=length_scale, win=roll_mean_win_size, alpha=alpha, limit=limit)
        self.t = 0.0
        lib.init(nx, ny, nz, lower, extent, rank, nranks, nccl_id)      # setup_domain_and_parameters
        lib.init_inversion(filtering)                                   # ps3d.f90:76

    def setup_fields(self, vor):
        """utils.f90:136-184."""
        o, lib = self.opts, self.lib
        lib.upload_vorticity(vor)
        lib.vor2vel()
        d = lib.diagnostics()
        self.nu = lib.init_diffusion(d["ke"], d["en"], o["nnu"], o["prediss"], o["length_scale"])
        lib.stepper_setup(o["stepper"])                                  # ps3d.f90:90-101
        return d

    def advance(self):
        o = self.opts
        self.t, dt, diag = self.lib.advance(self.t, o["limit"], o["alpha"], o["pretype"], o["win"])
        return dt, diag

    def close(self):
        self.lib.finalise()


def beltrami_solver(lib, n, **kw):
    """examples/beltrami_<n>.config: box [-pi/2, pi/2]^3, k = l = 2, m = 1."""
    lower = -0.5 * math.pi * np.ones(3)
    extent = math.pi * np.ones(3)
    s = Solver(lib, n, n, n, lower, extent, **kw)
    s.setup_fields(beltrami_vorticity(n, n, n, lower, extent))
    return s
