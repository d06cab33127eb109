"""Option coverage of the time-step path on the CPU block emulator against the oracle: every `pretype` of
advance.f90:385-408, both filters, molecular viscosity (nnu = 1), the 'geophysical' length scale, the
time-limit clip of dt (advance.f90:330-333), the rolling mean (rolling_mean.f90) and anisotropic grids."""
import math

import numpy as np
import pytest

import __graft_entry__ as G
from oracle import ps3d_oracle as O
from ps3d_b200.lib import PS3DLib

TOL = 5e-14


def rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.fixture(scope="module")
def emu():
    return PS3DLib(G.build_emu())


def run_pair(lib, shape, lower, extent, *, filtering="Hou & Li", nnu=3, prediss=30.0, length_scale="Kolmogorov",
             stepper="cn2", pretype="vorch", win=1000, limit=100.0, nsteps=2, seed=3, f_cor=None):
    nx, ny, nz = shape
    lib.init(nx, ny, nz, np.asarray(lower, float), np.asarray(extent, float))
    lib.init_inversion(filtering)
    try:
        s = O.PS3D(nx, ny, nz, lower, extent, filtering)
        if f_cor is not None:                       # planetary vorticity (physics.f90:163-169, inversion.f90:310-314)
            s.f_cor = np.asarray(f_cor, float)
            lib.set_physics(f_cor, 0.0)
        vor = np.random.default_rng(seed).uniform(-1, 1, (3, nx, ny, nz + 1))
        s.set_vorticity(vor, nnu=nnu, prediss=prediss, length_scale=length_scale)
        lib.upload_vorticity(vor)
        lib.vor2vel()
        d = lib.diagnostics()
        nu = lib.init_diffusion(d["ke"], d["en"], nnu, prediss, length_scale)
        assert nu == pytest.approx(s.vvisc, rel=1e-12)
        lib.stepper_setup(stepper)
        t = to = 0.0
        for i in range(nsteps):
            t, dt, diag = lib.advance(t, limit, 0.1, pretype, win)
            to, dto = s.advance(to, limit, stepper, 0.1, pretype, win, literal=True)
            assert dt == pytest.approx(dto, rel=1e-12), i
            assert diag["prefactor"] == pytest.approx(s.diag["pref"], rel=1e-11), i
            assert diag["rmv"] == pytest.approx(s.diag["rmv"], rel=1e-12), i
            assert rel(lib.download3("svor"), s.svor) < TOL, i
        return t, to
    finally:
        lib.finalise()


@pytest.mark.parametrize("pretype", ["constant", "vorch", "bfmax", "roll-mean-max-strain", "max-strain", "us-max-strain"])
def test_pretypes(emu, pretype):
    run_pair(emu, (8, 8, 8), [-0.5 * math.pi] * 3, [math.pi] * 3, pretype=pretype, win=2, nsteps=3)


@pytest.mark.parametrize("stepper", ["cn2", "impl-diff-rk4"])
def test_two_thirds_filter_and_molecular_viscosity(emu, stepper):
    run_pair(emu, (8, 16, 8), [0.0, 0.0, 0.0], [1.0, 2.0, 0.5], filtering="2/3-rule", nnu=1, prediss=2.0, stepper=stepper)


def test_geophysical_length_scale_and_anisotropic_grid(emu):
    run_pair(emu, (8, 32, 8), [-1.0, 0.0, -0.25], [2.0, 6.0, 0.5], length_scale="geophysical", prediss=10.0, nsteps=1)


def test_time_limit_clips_dt(emu):
    t, to = run_pair(emu, (8, 8, 8), [-0.5 * math.pi] * 3, [math.pi] * 3, limit=0.03, nsteps=2)
    assert t == pytest.approx(0.03, rel=1e-12) and to == pytest.approx(0.03, rel=1e-12)


@pytest.mark.parametrize("stepper", ["cn2", "impl-diff-rk4"])
def test_planetary_vorticity(emu, stepper):
    """f_cor != 0 without buoyancy: vor + f_cor in the tendency (inversion.f90:310-314), lat = 45 deg scaled up."""
    run_pair(emu, (8, 8, 8), [-0.5 * math.pi] * 3, [math.pi] * 3, stepper=stepper, f_cor=(0.0, 0.3, 0.3), nsteps=2)
