"""ENABLE_BUOYANCY code paths: the spectral diffz (inversion_utils.f90:683-719), buoyancy_tendency
(inversion.f90:232-292), r = u eta - v xi + b (:329-331), bfmax in adapt (advance.f90:147-168), the sbuoy updates of
both steppers (cn2.f90:107-117, 151-160; impl_rk4.f90:91-105, 124-131, 154-164, 187-195), the 'roll-mean-bfmax'
prefactor (advance.f90:395-398), planetary vorticity f_cor (inversion.f90:310-314) and the buoyancy pressure source
(fields_derived.f90:108-112).  The reference ships no test of these paths (`grep ENABLE_BUOYANCY unit-tests/` is
empty), so the oracle is pinned by an analytic known answer for diffz and the kernels by the oracle; the same cases
run on the CPU block emulator here and on the GPU in tests/test_gpu_buoyancy.py."""
import math

import numpy as np
import pytest

import __graft_entry__ as G
from oracle import ps3d_oracle as O
from ps3d_b200.lib import PS3DLib

TOL = 5e-13


def rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.fixture(scope="module")
def emu():
    return PS3DLib(G.build_emu())


def test_oracle_diffz_known_answer():
    """f = cos(2x) [sinh(2 z) + sin(3 pi (z - zlo) / Lz)]: a harmonic part (K = 2) plus one sine mode, for which the
    decomposition is exact, so combine(diffz(decompose(f))) = df/dz to round-off."""
    n = 32
    lower, extent = [-0.5 * math.pi] * 3, [math.pi] * 3
    s = O.PS3D(n, n, n, lower, extent)
    x = (lower[0] + extent[0] / n * np.arange(n))[:, None, None]
    z = (lower[2] + extent[2] / n * np.arange(n + 1))[None, None, :]
    m = 3.0 * math.pi / extent[2]
    f = np.cos(2 * x) * (np.sinh(2 * z) + np.sin(m * (z - lower[2]))) * np.ones((1, n, 1))
    dfdz = np.cos(2 * x) * (2 * np.cosh(2 * z) + m * np.cos(m * (z - lower[2]))) * np.ones((1, n, 1))
    got = s.field_combine_physical(s.diffz(s.field_decompose_physical(f)))
    assert rel(got, dfdz) < 1e-12


def _setup(lib, shape, lower, extent, *, stepper, f_cor, bfsq, bpretype, bwin, bnnu=3, seed=5):
    nx, ny, nz = shape
    lib.init(nx, ny, nz, np.asarray(lower, float), np.asarray(extent, float))
    lib.init_inversion("Hou & Li")
    s = O.PS3D(nx, ny, nz, lower, extent)
    rng = np.random.default_rng(seed)
    vor = rng.uniform(-1, 1, (3, nx, ny, nz + 1))
    buoy = rng.uniform(-0.5, 0.5, (nx, ny, nz + 1))
    s.f_cor = np.asarray(f_cor, float)
    s.set_vorticity(vor)
    s.enable_buoyancy(buoy, bfsq)
    lib.set_physics(f_cor, bfsq)
    lib.enable_buoyancy()
    lib.upload_vorticity(vor)
    lib.upload_buoyancy(buoy)
    lib.vor2vel()
    d = lib.diagnostics()
    lib.init_diffusion(d["ke"], d["en"])
    bv = lib.init_diffusion_buoyancy(d["ke"], d["en"], bnnu, 20.0, "Kolmogorov", bpretype, bwin)
    assert bv == pytest.approx(s.init_diffusion_buoyancy(d["ke"], d["en"], bnnu, 20.0), rel=1e-12)
    lib.stepper_setup(stepper)
    return s


def run_buoyancy(lib, shape, lower, extent, *, stepper="cn2", f_cor=(0.0, 1e-1, 2e-1), bfsq=0.7, pretype="vorch",
                 bpretype="roll-mean-bfmax", bwin=2, nsteps=3, bnnu=3):
    try:
        s = _setup(lib, shape, lower, extent, stepper=stepper, f_cor=f_cor, bfsq=bfsq, bpretype=bpretype, bwin=bwin, bnnu=bnnu)
        assert rel(lib.download("sbuoy"), s.sbuoy) < TOL
        t = to = 0.0
        for i in range(nsteps):
            t, dt, diag = lib.advance(t, 100.0, 0.1, pretype, 1000)
            to, dto = s.advance(to, 100.0, stepper, 0.1, pretype, 1000, literal=True, bpretype=bpretype, bwin=bwin)
            bd = lib.buoyancy_diag()
            assert dt == pytest.approx(dto, rel=1e-12), i
            assert diag["bfmax"] == pytest.approx(s.diag["bfmax"], rel=1e-12), i
            assert bd["rmb"] == pytest.approx(s.rmb, rel=1e-12), i
            assert bd["bval"] == pytest.approx(s.bpref, rel=1e-11), i
            assert rel(lib.download3("svor"), s.svor) < TOL, i
            assert rel(lib.download("sbuoy"), s.sbuoy) < TOL, i
        # the tendencies of the last source call and the output-only fields
        lib.vor2vel(); s.vor2vel()
        vor_rel = s.vor.copy()
        lib.source(); s.source()
        s.vor[:] = vor_rel                            # (the reference's source leaves the absolute vorticity in `vor`)
        assert rel(lib.download("sbuoys"), s.sbuoys) < TOL
        assert rel(lib.download3("svorts"), s.svorts) < TOL
        assert rel(lib.download("buoy"), s.field_combine_physical(s.sbuoy)) < TOL
        assert rel(lib.pressure(), s.pressure(*s.strain_fields())) < 1e-11
    finally:
        lib.finalise()


def run_diffz(lib, shape, lower, extent):
    nx, ny, nz = shape
    lib.init(nx, ny, nz, np.asarray(lower, float), np.asarray(extent, float))
    lib.init_inversion("Hou & Li")
    try:
        s = O.PS3D(nx, ny, nz, lower, extent)
        fs = s.field_decompose_physical(np.random.default_rng(1).uniform(-1, 1, (nx, ny, nz + 1)))
        assert rel(lib.diffz(fs), s.diffz(fs)) < TOL
    finally:
        lib.finalise()


def test_diffz_emu(emu):
    run_diffz(emu, (8, 16, 16), [-0.5 * math.pi] * 3, [math.pi, 2 * math.pi, 1.0])


@pytest.mark.parametrize("stepper", ["cn2", "impl-diff-rk4"])
def test_buoyancy_steps_emu(emu, stepper):
    run_buoyancy(emu, (8, 8, 8), [-0.5 * math.pi] * 3, [math.pi] * 3, stepper=stepper, nsteps=2)


def test_buoyancy_prefactors_emu(emu):
    run_buoyancy(emu, (8, 8, 8), [0.0, 0.0, 0.0], [2.0, 1.0, 0.5], pretype="bfmax", bpretype="vorch", nsteps=2, bnnu=1)


def test_buoyancy_needs_enable(emu):
    emu.init(8, 8, 8, np.zeros(3), np.ones(3))
    emu.init_inversion()
    try:
        with pytest.raises(Exception):
            emu.upload_buoyancy(np.zeros((8, 8, 9)))
    finally:
        emu.finalise()


def test_buoyancy_non_power_of_two_grid(emu):
    """nz = 12, ny = 12: the mixed-radix coverage kernels (k_zop_gen's spectral diffz, generic sweeps and updates)."""
    run_diffz(emu, (8, 12, 12), [-0.5 * math.pi] * 3, [math.pi, 2 * math.pi, 1.0])
    run_buoyancy(emu, (8, 12, 12), [-0.5 * math.pi] * 3, [math.pi] * 3, stepper="cn2", nsteps=2)
