"""Generates tests/golden/*.npz from the oracle (oracle/ps3d_oracle.py).

The reference ships no golden vectors and cannot be built in this image (no gfortran / MPI / netCDF), so these
are ORACLE-generated regression vectors: they pin the oracle (and through it the CUDA path) against drift, they
are not outputs of the Fortran build.  Run from the repo root:  python tests/golden/make_golden.py
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ps3d_oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def trajectory(stepper, n=32, nsteps=100):
    s = O.beltrami_setup(n)
    t = 0.0
    rows = []
    for _ in range(nsteps):
        t, dt = s.advance(t, 100.0, stepper, literal=True)
        rows.append([t, dt, s.diag["vortmax"], s.diag["vortrms"], s.diag["vorch"], s.diag["ggmax"], s.diag["umax"],
                     s.diag["vmax"], s.diag["wmax"], s.diag["usggmax"], s.diag["lsggmax"]])
    s.vor2vel()
    final = np.array([s.get_kinetic_energy(), s.get_enstrophy(), s.get_helicity()])
    # a thin sample of the final spectral vorticity (every 4th mode / level) keeps the file small
    return np.array(rows), final, s.svor[:, ::4, ::4, ::4].copy(), float(np.max(np.abs(s.svor)))


def operators(nx=16, ny=32, nz=8):
    lower = np.array([-math.pi, -0.5 * math.pi, 0.0])
    extent = np.array([2 * math.pi, math.pi, 1.7])
    s = O.PS3D(nx, ny, nz, lower, extent)
    f = np.random.default_rng(1234).uniform(-1, 1, (nx, ny, nz + 1))
    return dict(lower=lower, extent=extent, f=f, fftxyp2s=s.fftxyp2s(f), fftsine=s.fftsine(f), fftcosine=s.fftcosine(f),
                diffx=s.diffx(f), diffy=s.diffy(f), diffz=s.central_diffz(f),
                combine=s.field_combine_semi_spectral(f), decompose=s.field_decompose_semi_spectral(f))


if __name__ == "__main__":
    for stepper, tag in (("cn2", "cn2"), ("impl-diff-rk4", "rk4")):
        rows, final, sample, smax = trajectory(stepper)
        np.savez_compressed(os.path.join(OUT, f"beltrami32_{tag}_100steps.npz"), series=rows, final=final, svor_sample=sample,
                            svor_max=smax)
    np.savez_compressed(os.path.join(OUT, "operators_16x32x8.npz"), **operators())
    print("golden vectors written to", OUT)
