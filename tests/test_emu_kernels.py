"""Kernel index arithmetic on the CPU block emulator (tests/_emu, test-only build of the
same .cu sources with g++) against the oracle.  Small grids only; the parity tests proper
are the -m gpu ones."""
import math

import numpy as np
import pytest

import __graft_entry__ as G
from oracle import ps3d_oracle as O
from ps3d_b200.lib import PS3DLib, PS3DError

TOL = 5e-14


def rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.fixture(scope="module")
def emu():
    lib = PS3DLib(G.build_emu())
    yield lib


@pytest.fixture
def grid(emu):
    nx, ny, nz = 16, 32, 16
    lower = np.array([-0.5 * math.pi, -0.5 * math.pi, 0.0])
    extent = np.array([math.pi, 2 * math.pi, 1.5])
    emu.init(nx, ny, nz, lower, extent)
    emu.init_inversion("Hou & Li")
    yield emu, O.PS3D(nx, ny, nz, lower, extent), np.random.default_rng(1234)
    emu.finalise()


def test_operators(grid):
    lib, s, rng = grid
    f = rng.uniform(-1, 1, (s.nx, s.ny, s.nz + 1))
    fs = lib.fftxyp2s(f)
    assert rel(fs, s.fftxyp2s(f)) < TOL
    assert rel(lib.fftxys2p(fs), f) < TOL
    assert rel(lib.fftsine(f), s.fftsine(f)) < TOL
    assert rel(lib.fftcosine(f), s.fftcosine(f)) < TOL
    assert rel(lib.diffx(f), s.diffx(f)) < TOL
    assert rel(lib.diffy(f), s.diffy(f)) < TOL
    assert rel(lib.central_diffz(f), s.central_diffz(f)) < TOL
    assert rel(lib.field_combine_semi_spectral(f), s.field_combine_semi_spectral(f)) < TOL
    assert rel(lib.field_decompose_semi_spectral(f), s.field_decompose_semi_spectral(f)) < TOL
    assert rel(lib.field_combine_physical(f), s.field_combine_physical(f)) < TOL
    assert rel(lib.field_decompose_physical(f), s.field_decompose_physical(f)) < TOL


@pytest.mark.parametrize("stepper", ["cn2", "impl-diff-rk4"])
def test_time_step(grid, stepper):
    lib, s, rng = grid
    vor = rng.uniform(-1, 1, (3, s.nx, s.ny, s.nz + 1))
    s.set_vorticity(vor)
    lib.upload_vorticity(vor)
    lib.vor2vel()
    for name in ("svor", "vor", "svel", "vel"):
        assert rel(lib.download3(name), getattr(s, name)) < TOL, name
    d = lib.diagnostics()
    assert d["ke"] == pytest.approx(s.get_kinetic_energy(), rel=1e-13)
    assert d["en"] == pytest.approx(s.get_enstrophy(), rel=1e-13)
    assert d["helicity"] == pytest.approx(s.get_helicity(), rel=1e-11, abs=1e-16)
    hke = 0.5 * s._trap(s.vel[0] ** 2 + s.vel[1] ** 2) * s.ncelli           # field_diagnostics.f90:128-148
    hen = 0.5 * s._trap(s.vor[0] ** 2 + s.vor[1] ** 2) * s.ncelli           # :211-228
    assert d["hke"] == pytest.approx(hke, rel=1e-13) and d["vke"] == pytest.approx(s.get_kinetic_energy() - hke, rel=1e-11)
    assert d["hen"] == pytest.approx(hen, rel=1e-13) and d["ven"] == pytest.approx(s.get_enstrophy() - hen, rel=1e-11)
    assert d["hemax"] == pytest.approx(np.sqrt(np.max(s.vor[0] ** 2 + s.vor[1] ** 2)), rel=1e-13)
    lib.source()
    s.source()
    assert rel(lib.download3("svorts"), s.svorts) < TOL
    # output-only fields of adapt, evaluated lazily (fields_derived.f90:67-182)
    assert rel(lib.pressure(), s.pressure(*s.strain_fields())) < TOL
    assert rel(lib.horizontal_divergence(), s.horizontal_divergence()) < TOL
    lib.init_diffusion(d["ke"], d["en"])
    lib.stepper_setup(stepper)
    t, dt, diag = lib.advance(0.0, 100.0)
    to, dto = s.advance(0.0, 100.0, stepper, literal=True)
    assert dt == pytest.approx(dto, rel=1e-13) and t == pytest.approx(to, rel=1e-13)
    for k in ("vortmax", "vortrms", "vorch", "ggmax", "umax", "vmax", "wmax", "usggmax", "lsggmax"):
        assert diag[k] == pytest.approx(s.diag[k], rel=1e-12), k
    assert rel(lib.download3("svor"), s.svor) < TOL
    if stepper == "cn2":
        # the last tendency of the step (not stored by the time loop when the update rides on the source kernel:
        # re-evaluated on demand)
        assert rel(lib.download3("svorts"), s.svorts) < 1e-11


@pytest.mark.parametrize("stepper", ["cn2", "impl-diff-rk4"])
def test_separate_entry_points_equal_advance(grid, stepper):
    """vor2vel / adapt / source / step called one by one (the reference's module procedures) give the same step as
    ps3d_cuda_advance (which folds the first stepper update into the source kernel)."""
    lib, s, rng = grid
    vor = rng.uniform(-1, 1, (3, s.nx, s.ny, s.nz + 1))
    s.set_vorticity(vor)
    lib.upload_vorticity(vor)
    lib.vor2vel()
    d = lib.diagnostics()
    lib.init_diffusion(d["ke"], d["en"])
    lib.stepper_setup(stepper)
    lib.vor2vel()                                   # advance.f90:85
    dt, diag = lib.adapt(0.0, 100.0)                # :88 (sets the diffusion operator)
    lib.source()                                    # :95
    t = lib.step(0.0, dt)                           # :102
    to, dto = s.advance(0.0, 100.0, stepper, literal=True)
    assert dt == pytest.approx(dto, rel=1e-13) and t == pytest.approx(to, rel=1e-13)
    assert rel(lib.download3("svor"), s.svor) < TOL


def check_field_stats(lib, s):
    """All 40 scalars of the field-statistics file against the oracle (vor2vel + adapt done on both sides)."""
    got, want = lib.field_stats(), s.field_stats()
    assert set(got) == set(want) and len(got) == 40
    for k, w in want.items():
        g = got[k]
        if k in ("romin", "romax"):             # zeta / f_cor(3) with f_cor = 0: the same +-inf / nan as the reference
            assert (np.isnan(g) and np.isnan(w)) or g == w, k
        else:
            assert g == pytest.approx(w, rel=1e-11, abs=1e-15), k


def test_field_stats(grid):
    lib, s, rng = grid
    vor = rng.uniform(-1, 1, (3, s.nx, s.ny, s.nz + 1))
    s.set_vorticity(vor)
    lib.upload_vorticity(vor)
    lib.vor2vel()
    with pytest.raises(PS3DError):
        lib.field_stats()                        # needs the diagnostics of adapt
    d = lib.diagnostics()
    lib.init_diffusion(d["ke"], d["en"])
    lib.stepper_setup("cn2")
    lib.adapt(0.0, 100.0)
    s.adapt(0.0, 100.0)
    check_field_stats(lib, s)
    # ... and after a step, with the state the time loop leaves behind (write_step, advance.f90:92)
    t, dt, _ = lib.advance(0.0, 100.0)
    s.advance(0.0, 100.0, "cn2", literal=True)
    lib.vor2vel(); lib.adapt(t, 100.0)
    s.vor2vel(); s.adapt(t, 100.0)
    check_field_stats(lib, s)


@pytest.mark.parametrize("shape", [(24, 20, 12), (6, 6, 6), (16, 30, 16), (16, 16, 50)])
def test_non_power_of_two_grids(emu, shape):
    """Lengths 2^a 3^b 5^c that are not powers of two (stafft.f90:128-187 factorisen; radix 3 / 5 / 6 kernels
    :561-1757): mixed-radix coverage kernels (gen_fft.cuh, line_gen.cuh, zcol_gen.cuh), any mix of axes."""
    nx, ny, nz = shape
    lower = np.array([-0.5 * math.pi, 0.0, -1.0])
    extent = np.array([math.pi, 2 * math.pi, 2.0])
    emu.init(nx, ny, nz, lower, extent)
    try:
        emu.init_inversion("Hou & Li")
        s = O.PS3D(nx, ny, nz, lower, extent)
        rng = np.random.default_rng(3)
        f = rng.uniform(-1, 1, (nx, ny, nz + 1))
        fs = emu.fftxyp2s(f)
        assert rel(fs, s.fftxyp2s(f)) < TOL
        assert rel(emu.fftxys2p(fs), f) < TOL
        assert rel(emu.fftsine(f), s.fftsine(f)) < TOL
        assert rel(emu.fftcosine(f), s.fftcosine(f)) < TOL
        assert rel(emu.diffx(f), s.diffx(f)) < TOL
        assert rel(emu.diffy(f), s.diffy(f)) < TOL
        assert rel(emu.field_combine_physical(f), s.field_combine_physical(f)) < TOL
        assert rel(emu.field_decompose_physical(f), s.field_decompose_physical(f)) < TOL
        vor = rng.uniform(-1, 1, (3, nx, ny, nz + 1))
        s.set_vorticity(vor)
        emu.upload_vorticity(vor)
        emu.vor2vel()
        for name in ("svor", "vor", "svel", "vel"):
            assert rel(emu.download3(name), getattr(s, name)) < TOL, name
        d = emu.diagnostics()
        emu.init_diffusion(d["ke"], d["en"])
        emu.stepper_setup("cn2")
        t, dt, _ = emu.advance(0.0, 100.0)
        to, dto = s.advance(0.0, 100.0, "cn2", literal=True)
        assert dt == pytest.approx(dto, rel=1e-13)
        assert rel(emu.download3("svor"), s.svor) < TOL
    finally:
        emu.finalise()


def check_reference_z_known_answers(lib, nz_deriv):
    """The reference's analytic z tests through the C ABI: unit-tests/test_diffz_1..4.f90 (central_diffz) and
    tests/test_deriv.f90 (fftsine -> * rkz -> fftcosine), with the reference's tolerances."""
    from test_oracle_known_answers import deriv_via_sine_series, grid as ogrid
    PI = math.pi
    cases = [((32, 32, 32), [0.0, 0.0, 0.0], [1.0, 1.0, 1.0], lambda x, y, z: z + 0 * x * y, lambda x, y, z: 1.0 + 0 * z, 1e-14),
             ((16, 32, 32), [0.0, -0.5 * PI, 0.0], [PI, 2 * PI, 2 * PI], lambda x, y, z: z * np.cos(2 * x) * np.sin(y),
              lambda x, y, z: np.cos(2 * x) * np.sin(y) + 0 * z, 2e-14),
             ((16, 32, 32), [0.0, -0.5 * PI, 0.0], [PI, 2 * PI, 2 * PI], lambda x, y, z: 6 * z + 3 * np.cos(2 * x) * np.sin(y) * z,
              lambda x, y, z: 6 + 3 * np.cos(2 * x) * np.sin(y) + 0 * z, 2e-13),
             ((32, 32, 32), [0.0, 0.0, 0.0], [1.0, 1.0, 1.0], lambda x, y, z: 1 - 3 * z ** 2 + 0 * x * y, lambda x, y, z: -6 * z + 0 * x * y, 1e-1)]
    for (nx, ny, nz), lower, extent, f, dfdz, atol in cases:
        lib.init(nx, ny, nz, np.asarray(lower, float), np.asarray(extent, float))
        try:
            lib.init_inversion("Hou & Li")
            x, y, z = ogrid(O.PS3D(nx, ny, nz, lower, extent))
            assert np.max(np.abs(lib.central_diffz(f(x, y, z)) - dfdz(x, y, z))) < atol
        finally:
            lib.finalise()
    nz = nz_deriv
    lib.init(8, 8, nz, np.zeros(3), np.ones(3))
    try:
        lib.init_inversion("Hou & Li")
        col = lambda v: np.broadcast_to(v, (8, 8, nz + 1)).copy()
        d, exact = deriv_via_sine_series(nz, lambda a: lib.fftsine(col(a))[3, 5], lambda v: lib.fftcosine(col(v))[3, 5])
        assert np.max(np.abs(d - exact)) < 3.01 / nz
        want, _ = deriv_via_sine_series(nz, lambda a: np.concatenate(([0.0], O.dst(a[1:].copy(), nz))), lambda v: O.dct(v, nz))
        # The derivative is cosine(rkz * sine(f)): absolute round-off of the sine coefficients (~1e-16 x their largest
        # partial sum, uniform in kz) is amplified by rkz <= pi nz.  Two bars, both tied to the reference: relative
        # to the magnitude the transforms carry, max |rkz * w| (deriv1d.f90:17-22), and the round-off floor of the
        # reference's OWN arithmetic for this very test (literal stafft.f90 dst -> rkz -> dct, oracle/stafft_lit.c:
        # 1.7e-10 at nz = 1024, 2.7e-13 at nz = 64).  The CUDA transforms must be at least as accurate as the Fortran.
        import __graft_entry__ as G
        G.build_stafft_lit()
        from oracle.stafft_lit import Stafft
        S = Stafft(nz)
        lit, _ = deriv_via_sine_series(nz, lambda a: np.concatenate(([0.0], S.dst(a[1:].copy()))), lambda v: S.dct(v))
        z = np.arange(nz + 1) / nz
        w = np.concatenate(([0.0], O.dst(((1 - 2 * z ** 3) - (1 - 2 * z))[1:].copy(), nz)))
        amp = float(np.max(np.abs(math.pi * np.arange(nz + 1) * w)))
        floor = float(np.max(np.abs(lit - want)))
        err = float(np.max(np.abs(d - want)))
        assert err < max(1e-12 * amp, floor), (err, amp, floor)
    finally:
        lib.finalise()


def test_reference_z_known_answers(emu):
    check_reference_z_known_answers(emu, 64)


def check_genspec(lib, s):
    """Kinetic-energy spectrum (genspec.f90:55-127) against the oracle; sum(spec) dk = kinetic energy."""
    spec, num, dk = lib.genspec()
    wspec, wnum, wdk = s.genspec()
    assert dk == pytest.approx(wdk, rel=1e-14) and len(spec) == len(wspec)
    assert np.array_equal(num, wnum)
    assert np.max(np.abs(spec - wspec)) < 1e-12 * np.max(wspec)
    assert np.sum(spec) * dk == pytest.approx(s.get_kinetic_energy(), rel=1e-12)


def test_genspec(grid):
    lib, s, rng = grid
    vor = rng.uniform(-1, 1, (3, s.nx, s.ny, s.nz + 1))
    s.set_vorticity(vor)
    lib.upload_vorticity(vor)
    lib.vor2vel()
    check_genspec(lib, s)


def test_error_paths(emu):
    with pytest.raises(PS3DError) as e:
        emu.vor2vel()
    assert e.value.status == 1
    with pytest.raises(PS3DError) as e:
        emu.init(28, 32, 32, np.zeros(3), np.ones(3))       # a factor 7: stafft.f90:87-95
    assert e.value.status == 3
    with pytest.raises(PS3DError) as e:
        emu.init(32, 32, 32, np.zeros(3), np.array([1.0, 0.0, 1.0]))   # sta2dfft.f90:67-74
    assert e.value.status == 2


def test_genspec_large_box(emu):
    """on a box larger than pi (2, 2, 1) the shell index int(kmag / dk) exceeds kmax (the reference
    allocates spec(0:kmax), genspec.f90:86-100: latent overflow).  Every point must land in a bin and the bins
    must integrate to the kinetic energy."""
    n = 16
    lower, extent = -math.pi * np.ones(3), 2 * math.pi * np.ones(3)
    emu.init(n, n, n, lower, extent)
    try:
        emu.init_inversion("Hou & Li")
        s = O.PS3D(n, n, n, lower, extent)
        vor = np.random.default_rng(2).uniform(-1, 1, (3, n, n, n + 1))
        s.set_vorticity(vor)
        emu.upload_vorticity(vor)
        emu.vor2vel()
        spec, num, dk = emu.genspec()
        assert dk < 1.0 and len(num) > int(round(math.sqrt(3 * (n / 2) ** 2))) + 1
        assert np.sum(num) == n * n * (n + 1)
        check_genspec(emu, s)
        # the state next to the bins is intact: a second vor2vel reproduces the oracle's second vor2vel
        emu.vor2vel()
        s.vor2vel()
        assert rel(emu.download3("vel"), s.vel) < TOL
    finally:
        emu.finalise()


def test_rk4_two_steps_without_new_diffusion(grid):
    """ps3d_cuda_step is the drop-in for bstep%step; impl_rk4_step rebuilds epq / emq from vdiss at every
    call (impl_rk4.f90:87-89), so two steps after ONE set_diffusion must not see squared factors."""
    lib, s, rng = grid
    vor = rng.uniform(-1, 1, (3, s.nx, s.ny, s.nz + 1))
    s.set_vorticity(vor)
    lib.upload_vorticity(vor)
    lib.vor2vel()
    d = lib.diagnostics()
    lib.init_diffusion(d["ke"], d["en"])
    lib.stepper_setup("impl-diff-rk4")
    dt, _ = lib.adapt(0.0, 100.0)
    dto, pref = s.adapt(0.0, 100.0)              # (set_vorticity has run the oracle's vor2vel)
    s.rk4_set_diffusion(dto, pref)
    t = to = 0.0
    for i in range(2):
        lib.source()
        t = lib.step(t, dt)
        s.source()
        to = s.rk4_step(to, dto, literal=True)
        assert rel(lib.download3("svor"), s.svor) < TOL, i
        lib.vor2vel()
        s.vor2vel()


def test_rank_limit_and_window_change(emu):
    with pytest.raises(PS3DError) as e:
        emu.init(32, 32, 8, np.zeros(3), np.ones(3), rank=0, nranks=16)
    assert e.value.status == 2
    emu.init(8, 8, 8, np.zeros(3), np.ones(3))
    try:
        emu.init_inversion("Hou & Li")
        vor = np.random.default_rng(1).uniform(-1, 1, (3, 8, 8, 9))
        emu.upload_vorticity(vor)
        emu.vor2vel()
        d = emu.diagnostics()
        emu.init_diffusion(d["ke"], d["en"])
        emu.stepper_setup("cn2")
        emu.adapt(0.0, 100.0, win=4)
        with pytest.raises(PS3DError) as e:
            emu.adapt(0.0, 100.0, win=64)          # rolling_mean.f90: the window is allocated once
        assert e.value.status == 2
    finally:
        emu.finalise()


def test_streamed_upload_matches_blocking_upload(emu):
    """ps3d_cuda_upload_vorticity_begin/_end leave the same state as ps3d_cuda_upload_vorticity, and the pending
    copy does not disturb the state it will replace."""
    n = 8
    emu.init(n, n, n, np.zeros(3), np.ones(3))
    emu.init_inversion()
    try:
        rng = np.random.default_rng(11)
        v1, v2 = rng.uniform(-1, 1, (3, n, n, n + 1)), rng.uniform(-1, 1, (3, n, n, n + 1))
        emu.upload_vorticity(v1)
        s1 = emu.download3("svor")
        emu.upload_vorticity(v2)
        s2 = emu.download3("svor")
        emu.upload_vorticity(v1)
        emu.upload_vorticity_begin(v2)
        assert np.array_equal(emu.download3("svor"), s1)
        emu.upload_vorticity_begin(v1)                 # a second upload may be queued (two staging sets) ...
        with pytest.raises(Exception):
            emu.upload_vorticity_begin(v2)             # ... a third may not
        emu.upload_vorticity_end()
        assert np.array_equal(emu.download3("svor"), s2)         # first in, first out
        emu.upload_vorticity_end()
        assert np.array_equal(emu.download3("svor"), s1)
        with pytest.raises(Exception):
            emu.upload_vorticity_end()
    finally:
        emu.finalise()


def test_step_needs_a_source_call(emu):
    """bstep%step consumes the tendency of the current state (advance.f90:95-102): after ps3d_cuda_advance (whose
    source kernel carried the update and kept no svorts) a bare ps3d_cuda_step is refused; with a source call it runs."""
    from ps3d_b200 import host
    s = host.beltrami_solver(emu, 8, stepper="cn2")
    try:
        dt, _ = s.advance()
        with pytest.raises(Exception):
            emu.step(s.t, dt)
        emu.vor2vel(); emu.source()
        emu.step(s.t, dt)
    finally:
        s.close()
