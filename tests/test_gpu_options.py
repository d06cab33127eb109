"""Option coverage of the time-step path ON THE GPU through the C ABI against the oracle (the same cases
tests/test_emu_options.py runs on the CPU block emulator): every `pretype` of advance.f90:385-408, both filters,
molecular viscosity (nnu = 1), the 'geophysical' length scale (inversion_utils.f90:157-184), the time-limit clip of
dt (advance.f90:330-333), the rolling mean (rolling_mean.f90) and anisotropic boxes / grids."""
import math

import pytest

from test_emu_options import run_pair

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    import torch
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    import ps3d_b200
    return ps3d_b200.load()


@pytest.mark.parametrize("pretype", ["constant", "vorch", "bfmax", "roll-mean-max-strain", "max-strain", "us-max-strain"])
def test_pretypes(lib, pretype):
    run_pair(lib, (16, 16, 16), [-0.5 * math.pi] * 3, [math.pi] * 3, pretype=pretype, win=2, nsteps=3)


@pytest.mark.parametrize("stepper", ["cn2", "impl-diff-rk4"])
def test_two_thirds_filter_and_molecular_viscosity(lib, stepper):
    run_pair(lib, (16, 32, 16), [0.0, 0.0, 0.0], [1.0, 2.0, 0.5], filtering="2/3-rule", nnu=1, prediss=2.0, stepper=stepper)


def test_geophysical_length_scale_and_anisotropic_grid(lib):
    run_pair(lib, (16, 64, 32), [-1.0, 0.0, -0.25], [2.0, 6.0, 0.5], length_scale="geophysical", prediss=10.0, nsteps=2)


def test_time_limit_clips_dt(lib):
    t, to = run_pair(lib, (16, 16, 16), [-0.5 * math.pi] * 3, [math.pi] * 3, limit=0.03, nsteps=2)
    assert t == pytest.approx(0.03, rel=1e-12) and to == pytest.approx(0.03, rel=1e-12)


@pytest.mark.parametrize("stepper", ["cn2", "impl-diff-rk4"])
def test_planetary_vorticity(lib, stepper):
    run_pair(lib, (32, 16, 32), [-0.5 * math.pi] * 3, [math.pi] * 3, stepper=stepper, f_cor=(0.0, 0.3, 0.3), nsteps=2)
