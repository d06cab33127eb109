"""The ENABLE_BUOYANCY code paths ON THE GPU through the C ABI against the oracle (the cases of
tests/test_buoyancy.py at sizes that use the register-blocked kernels): spectral diffz (inversion_utils.f90:683-719),
buoyancy_tendency (inversion.f90:232-292), bfmax (advance.f90:147-168), sbuoy stepping with cn2 and impl-diff-rk4,
'roll-mean-bfmax', planetary vorticity and the buoyancy pressure source (fields_derived.f90:108-112)."""
import math

import pytest

from test_buoyancy import run_buoyancy, run_diffz

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    import torch
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    import ps3d_b200
    return ps3d_b200.load()


@pytest.mark.parametrize("shape", [(16, 16, 16), (32, 8, 64), (8, 16, 256), (8, 8, 512)])
def test_diffz(lib, shape):
    run_diffz(lib, shape, [-0.5 * math.pi] * 3, [math.pi, 2 * math.pi, 1.0])


@pytest.mark.parametrize("stepper", ["cn2", "impl-diff-rk4"])
def test_buoyancy_steps(lib, stepper):
    run_buoyancy(lib, (32, 32, 32), [-0.5 * math.pi] * 3, [math.pi] * 3, stepper=stepper, nsteps=3)


def test_buoyancy_prefactors_and_molecular_diffusion(lib):
    run_buoyancy(lib, (16, 32, 16), [0.0, 0.0, 0.0], [2.0, 1.0, 0.5], pretype="bfmax", bpretype="vorch", nsteps=2, bnnu=1)


def test_buoyancy_64(lib):
    run_buoyancy(lib, (64, 64, 64), [-0.5 * math.pi] * 3, [math.pi] * 3, stepper="cn2", nsteps=2, f_cor=(0.0, 0.0, 0.0), bfsq=0.0)


def test_buoyancy_non_power_of_two_grid(lib):
    run_diffz(lib, (12, 24, 36), [-0.5 * math.pi] * 3, [math.pi, 2 * math.pi, 1.0])
    run_buoyancy(lib, (24, 12, 36), [-0.5 * math.pi] * 3, [math.pi] * 3, stepper="cn2", nsteps=2)
