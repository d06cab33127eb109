"""The C++/OpenMP restatement of the reference path (oracle/ps3d_ref.cpp: the CPU baseline of bench.py, with the
reference's transposes, reversed copies, stored tables and literal stepper) against the NumPy oracle."""
import math

import numpy as np
import pytest

import __graft_entry__ as G
from oracle import ps3d_oracle as O
from oracle.ps3d_ref import RefSolver, OPS

PI = math.pi


def rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize("shape,filtering,stepper", [((16, 32, 8), "Hou & Li", "cn2"), ((32, 16, 32), "2/3-rule", "cn2"),
                                                     ((16, 16, 16), "Hou & Li", "impl-diff-rk4")])
def test_cpp_restatement_matches_numpy_oracle(shape, filtering, stepper):
    G.build_ref()
    nx, ny, nz = shape
    lower = np.array([-0.5 * PI, 0.0, -1.0])
    extent = np.array([PI, 2 * PI, 2.0])
    r = RefSolver(nx, ny, nz, lower, extent, filtering)
    s = O.PS3D(nx, ny, nz, lower, extent, filtering)
    try:
        rng = np.random.default_rng(1)
        f = rng.uniform(-1, 1, (nx, ny, nz + 1))
        for name in OPS:
            assert rel(r.op(name, f), getattr(s, name)(f)) < 1e-13, name
        assert rel(r.op("fftxys2p", r.op("fftxyp2s", f)), f) < 1e-13
        vor = rng.uniform(-1, 1, (3, nx, ny, nz + 1))
        ke, en = r.set_vorticity(vor)
        ke0, en0 = s.set_vorticity(vor)
        assert ke == pytest.approx(ke0, rel=1e-13) and en == pytest.approx(en0, rel=1e-13)
        for name in ("svor", "vor", "vel", "svel"):
            assert rel(r.get(name), getattr(s, name)) < 1e-13, name
        t0 = 0.0
        for _ in range(2):
            t, dt = r.advance(stepper=stepper)
            t0, dt0 = s.advance(t0, 100.0, stepper, literal=True)
            assert dt == pytest.approx(dt0, rel=1e-12) and t == pytest.approx(t0, rel=1e-12)
            assert r.diag()["vorch"] == pytest.approx(s.diag["vorch"], rel=1e-12)
            assert r.diag()["ggmax"] == pytest.approx(s.diag["ggmax"], rel=1e-12)
            assert rel(r.get("svor"), s.svor) < 1e-12
            if stepper == "cn2":
                assert rel(r.get("svorts"), s.svorts) < 1e-11
    finally:
        r.close()


def test_cpp_restatement_beltrami_known_answer():
    """unit-tests/test_vor2vel_1.f90: Beltrami flow 32^3, velocity = vorticity / alpha-type known answer through the
    restatement: vor2vel reproduces the analytic velocity to 1e-13."""
    G.build_ref()
    n = 32
    lower = -0.5 * PI * np.ones(3)
    extent = PI * np.ones(3)
    r = RefSolver(n, n, n, lower, extent)
    try:
        vor = O.beltrami_vorticity(n, n, n, lower, extent)
        r.set_vorticity(vor)
        alpha = 3.0                      # sqrt(k^2 + l^2 + m^2), k = l = 2, m = 1: u = omega / alpha
        assert np.max(np.abs(r.get("vel") - vor / alpha)) < 1e-13
    finally:
        r.close()


def test_two_restatements_agree_over_100_steps():
    """Beltrami 32^3, 100 cn2 steps: the NumPy oracle and the independent C++ restatement
    (different FFT code, different sweep structure) give the same dt sequence and state -- the trajectory the CUDA
    path is compared with is not an artefact of one restatement."""
    G.build_ref()
    n = 32
    lower = -0.5 * PI * np.ones(3)
    extent = PI * np.ones(3)
    r = RefSolver(n, n, n, lower, extent)
    s = O.beltrami_setup(n)
    try:
        r.set_vorticity(O.beltrami_vorticity(n, n, n, lower, extent))
        t0 = 0.0
        worst_dt = 0.0
        for _ in range(100):
            t, dt = r.advance()
            t0, dt0 = s.advance(t0, 100.0, "cn2", literal=True)
            worst_dt = max(worst_dt, abs(dt - dt0) / dt0)
        assert worst_dt < 1e-11
        assert t == pytest.approx(t0, rel=1e-12)
        assert rel(r.get("svor"), s.svor) < 1e-11
        vel, vor = r.get("vel"), r.get("vor")
        ke = 0.5 * s._trap(vel[0] ** 2 + vel[1] ** 2 + vel[2] ** 2) * s.ncelli
        assert ke == pytest.approx(s.get_kinetic_energy(), rel=1e-10)
    finally:
        r.close()
