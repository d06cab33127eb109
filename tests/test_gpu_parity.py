"""Parity tests proper: the CUDA path through the C ABI against the oracle on the same inputs
(bar from BASELINE.json north_star: per-step fields <= 1e-12 relative to the field max-norm,
energy / enstrophy / helicity <= 1e-10 relative after 100 steps), the reference's own analytic
known-answer tests through the C ABI, and size-independent properties at full size."""
import math

import numpy as np
import pytest

from oracle import ps3d_oracle as O

pytestmark = pytest.mark.gpu

PI = math.pi
FIELD_TOL = 1e-12      # north_star: per-step fields, relative (max-norm)
DIAG_TOL = 1e-10       # north_star: KE / enstrophy / helicity after 100 steps


def rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.fixture(scope="module")
def lib():
    import torch
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    import ps3d_b200
    return ps3d_b200.load()


def open_grid(lib, nx, ny, nz, lower, extent, filtering="Hou & Li"):
    lib.init(nx, ny, nz, np.asarray(lower, float), np.asarray(extent, float))
    lib.init_inversion(filtering)
    return O.PS3D(nx, ny, nz, lower, extent, filtering)


@pytest.mark.parametrize("shape", [(64, 64, 64), (16, 32, 8), (128, 64, 32), (8, 8, 8), (16, 16, 1024), (1024, 8, 16),
                                   # the kernel instantiations of the benchmarked configurations (k_line_*<512>, <256> on
                                   # both axes, full-width nz = 512 columns)
                                   (512, 512, 8), (512, 8, 16), (8, 512, 16), (256, 256, 16), (64, 64, 512)])
def test_operators_white_noise(lib, shape):
    """isolated transforms on a white-noise field (all modes populated)."""
    nx, ny, nz = shape
    s = open_grid(lib, nx, ny, nz, [-PI, -0.5 * PI, 0.0], [2 * PI, PI, 1.7])
    try:
        f = np.random.default_rng(1234).uniform(-1, 1, (nx, ny, nz + 1))
        fs = lib.fftxyp2s(f)
        assert rel(fs, s.fftxyp2s(f)) < FIELD_TOL
        assert rel(lib.fftxys2p(fs), f) < FIELD_TOL
        assert rel(lib.fftsine(f), s.fftsine(f)) < FIELD_TOL
        assert rel(lib.fftcosine(f), s.fftcosine(f)) < FIELD_TOL
        assert rel(lib.diffx(f), s.diffx(f)) < FIELD_TOL
        assert rel(lib.diffy(f), s.diffy(f)) < FIELD_TOL
        assert rel(lib.central_diffz(f), s.central_diffz(f)) < FIELD_TOL
        assert rel(lib.field_combine_semi_spectral(f), s.field_combine_semi_spectral(f)) < FIELD_TOL
        assert rel(lib.field_decompose_semi_spectral(f), s.field_decompose_semi_spectral(f)) < FIELD_TOL
        assert rel(lib.field_combine_physical(f), s.field_combine_physical(f)) < FIELD_TOL
        assert rel(lib.field_decompose_physical(f), s.field_decompose_physical(f)) < FIELD_TOL
        # 2/3 rule shares everything but the filter tables
    finally:
        lib.finalise()


def test_vor2vel_1_known_answer(lib):
    """unit-tests/test_vor2vel_1.f90:99: Beltrami 32^3, velocity vs analytic, atol 1e-14."""
    s = open_grid(lib, 32, 32, 32, -0.5 * PI * np.ones(3), PI * np.ones(3))
    try:
        x = (s.lower[0] + s.dx[0] * np.arange(32))[:, None, None]
        y = (s.lower[1] + s.dx[1] * np.arange(32))[None, :, None]
        z = (s.lower[2] + s.dx[2] * np.arange(33))[None, None, :]
        k, l, m = 2.0, 2.0, 1.0
        alpha = math.sqrt(k * k + l * l + m * m)
        f = 1.0 / (k * k + l * l)
        ref = np.empty((3, 32, 32, 33))
        ref[0] = f * (k * m * np.sin(m * z) - l * alpha * np.cos(m * z)) * np.sin(k * x + l * y)
        ref[1] = f * (l * m * np.sin(m * z) + k * alpha * np.cos(m * z)) * np.sin(k * x + l * y)
        ref[2] = np.cos(m * z) * np.cos(k * x + l * y)
        lib.upload_vorticity(alpha * ref)
        lib.vor2vel()
        assert np.max(np.abs(lib.download3("vel") - ref)) < 1e-14
    finally:
        lib.finalise()


def test_vor2vel_mean_flow_known_answers(lib):
    """unit-tests/test_vor2vel_3.f90:82 (eta = 1 -> u = z - zc, 1e-15) and _4 (eta = 2z)."""
    s = open_grid(lib, 32, 32, 32, [-0.5, -0.5, 0.0], [1.0, 1.0, 1.0])
    try:
        z = (s.lower[2] + s.dx[2] * np.arange(33))[None, None, :] + np.zeros((32, 32, 1))
        vor = np.zeros((3, 32, 32, 33))
        vor[1] = 1.0
        lib.upload_vorticity(vor)
        lib.vor2vel()
        vel = lib.download3("vel")
        assert np.max(np.abs(vel[0] - (z - 0.5))) < 2e-15 and np.max(np.abs(vel[1:])) < 2e-15
        vor[1] = 2 * z
        lib.upload_vorticity(vor)
        lib.vor2vel()
        vel = lib.download3("vel")
        assert np.max(np.abs(vel[0] - (z ** 2 - 1.0 / 3.0))) < 2e-15
    finally:
        lib.finalise()


def test_diffx_diffy_known_answer(lib):
    """unit-tests/test_diffx.f90:59-75 / test_diffy.f90: cos 4x -> -4 sin 4x, 1e-12."""
    s = open_grid(lib, 64, 128, 64, -PI * np.ones(3), 2 * PI * np.ones(3))
    try:
        x = (s.lower[0] + s.dx[0] * np.arange(64))[:, None, None] + np.zeros((1, 128, 65))
        out = lib.fftxys2p(lib.diffx(lib.fftxyp2s(np.cos(4 * x))))
        assert np.max(np.abs(out + 4 * np.sin(4 * x))) < 1e-12
        y = (s.lower[1] + s.dx[1] * np.arange(128))[None, :, None] + np.zeros((64, 1, 65))
        out = lib.fftxys2p(lib.diffy(lib.fftxyp2s(np.cos(4 * y))))
        assert np.max(np.abs(out + 4 * np.sin(4 * y))) < 1e-12
    finally:
        lib.finalise()


@pytest.mark.parametrize("filtering", ["Hou & Li", "2/3-rule"])
def test_vor2vel_source_white_noise_64(lib, filtering):
    s = open_grid(lib, 64, 64, 64, -0.5 * PI * np.ones(3), PI * np.ones(3), filtering)
    try:
        vor = np.random.default_rng(7).uniform(-1, 1, (3, 64, 64, 65))
        s.set_vorticity(vor)
        lib.upload_vorticity(vor)
        lib.vor2vel()
        for name in ("svor", "vor", "svel", "vel"):
            assert rel(lib.download3(name), getattr(s, name)) < FIELD_TOL, name
        lib.source()
        s.source()
        assert rel(lib.download3("svorts"), s.svorts) < FIELD_TOL
        assert rel(lib.pressure(), s.pressure(*s.strain_fields())) < FIELD_TOL          # fields_derived.f90:67
        assert rel(lib.horizontal_divergence(), s.horizontal_divergence()) < FIELD_TOL   # fields_derived.f90:161
        d = lib.diagnostics()
        assert d["ke"] == pytest.approx(s.get_kinetic_energy(), rel=1e-13)
        assert d["en"] == pytest.approx(s.get_enstrophy(), rel=1e-13)
    finally:
        lib.finalise()


@pytest.mark.parametrize("shape", [(16, 16, 512), (8, 16, 256), (16, 8, 128), (8, 8, 1024), (32, 32, 16),
                                   (512, 512, 8), (512, 8, 16), (8, 512, 16), (256, 256, 16), (64, 64, 512)])
def test_vor2vel_source_white_noise_tall(lib, shape):
    """Column kernels at every nz template the configs use (512: three blocks per SM, in-place transforms; 1024),
    both instantiations (pairs (a, ny-a) and the (0, ny/2) pair), anisotropic boxes, against the oracle."""
    nx, ny, nz = shape
    s = open_grid(lib, nx, ny, nz, [-0.5 * PI, 0.0, -1.0], [PI, 2 * PI, 2.0])
    try:
        vor = np.random.default_rng(11).uniform(-1, 1, (3, nx, ny, nz + 1))
        s.set_vorticity(vor)
        lib.upload_vorticity(vor)
        lib.vor2vel()
        for name in ("svor", "vor", "svel", "vel"):
            assert rel(lib.download3(name), getattr(s, name)) < FIELD_TOL, name
        lib.source()
        s.source()
        assert rel(lib.download3("svorts"), s.svorts) < FIELD_TOL
    finally:
        lib.finalise()


@pytest.mark.parametrize("shape", [(48, 40, 36), (96, 96, 96), (12, 24, 384), (384, 12, 24), (16, 640, 10)])
def test_non_power_of_two_grids(lib, shape):
    """the lengths factorisen accepts beyond powers of two (radix 3 / 5 / 6, stafft.f90:128-187,
    561-1757) through the mixed-radix coverage kernels; operators, vor2vel, source and one cn2 step."""
    nx, ny, nz = shape
    s = open_grid(lib, nx, ny, nz, [-0.5 * PI, 0.0, -1.0], [PI, 2 * PI, 2.0])
    try:
        rng = np.random.default_rng(3)
        f = rng.uniform(-1, 1, (nx, ny, nz + 1))
        fs = lib.fftxyp2s(f)
        assert rel(fs, s.fftxyp2s(f)) < FIELD_TOL
        assert rel(lib.fftxys2p(fs), f) < FIELD_TOL
        assert rel(lib.fftsine(f), s.fftsine(f)) < FIELD_TOL
        assert rel(lib.fftcosine(f), s.fftcosine(f)) < FIELD_TOL
        assert rel(lib.diffx(f), s.diffx(f)) < FIELD_TOL
        assert rel(lib.diffy(f), s.diffy(f)) < FIELD_TOL
        assert rel(lib.field_decompose_physical(f), s.field_decompose_physical(f)) < FIELD_TOL
        vor = rng.uniform(-1, 1, (3, nx, ny, nz + 1))
        s.set_vorticity(vor)
        lib.upload_vorticity(vor)
        lib.vor2vel()
        for name in ("svor", "vor", "svel", "vel"):
            assert rel(lib.download3(name), getattr(s, name)) < FIELD_TOL, name
        lib.source()
        s.source()
        assert rel(lib.download3("svorts"), s.svorts) < FIELD_TOL
        d = lib.diagnostics()
        lib.init_diffusion(d["ke"], d["en"])
        lib.stepper_setup("cn2")
        t, dt, _ = lib.advance(0.0, 100.0)
        to, dto = s.advance(0.0, 100.0, "cn2", literal=True)
        assert dt == pytest.approx(dto, rel=1e-11)
        assert rel(lib.download3("svor"), s.svor) < FIELD_TOL
    finally:
        lib.finalise()


def test_reference_z_known_answers(lib):
    """unit-tests/test_diffz_1..4.f90 and tests/test_deriv.f90 (nz = 1024 as in the reference) through the C ABI."""
    from test_emu_kernels import check_reference_z_known_answers
    check_reference_z_known_answers(lib, 1024)


def test_genspec_64(lib):
    """the kinetic-energy spectrum of genspec.f90 as a device reduction."""
    from test_emu_kernels import check_genspec
    s = open_grid(lib, 64, 64, 64, -0.5 * PI * np.ones(3), PI * np.ones(3))
    try:
        vor = np.random.default_rng(9).uniform(-1, 1, (3, 64, 64, 65))
        s.set_vorticity(vor)
        lib.upload_vorticity(vor)
        lib.vor2vel()
        check_genspec(lib, s)
    finally:
        lib.finalise()


def test_field_stats_64(lib):
    """the 40 scalars of the field-statistics file (field_diagnostics_netcdf.f90:257-439)."""
    from test_emu_kernels import check_field_stats
    s = open_grid(lib, 64, 64, 64, -0.5 * PI * np.ones(3), PI * np.ones(3))
    try:
        vor = np.random.default_rng(5).uniform(-1, 1, (3, 64, 64, 65))
        s.set_vorticity(vor)
        lib.upload_vorticity(vor)
        lib.vor2vel()
        d = lib.diagnostics()
        lib.init_diffusion(d["ke"], d["en"])
        lib.stepper_setup("cn2")
        lib.adapt(0.0, 100.0)
        s.adapt(0.0, 100.0)
        check_field_stats(lib, s)
    finally:
        lib.finalise()


@pytest.mark.parametrize("stepper,n,nsteps", [("cn2", 32, 100), ("impl-diff-rk4", 32, 100), ("cn2", 64, 10)])
def test_beltrami_trajectory(lib, stepper, n, nsteps):
    """fields every step to 1e-12, dt sequence, and
    KE / enstrophy / helicity after the last step to 1e-10 (the oracle runs the *literal*
    combine -> multiply -> decompose steppers of the reference)."""
    from ps3d_b200 import host
    ref = O.beltrami_setup(n)
    s = host.beltrami_solver(lib, n, stepper=stepper)
    try:
        t = 0.0
        for i in range(nsteps):
            dt, diag = s.advance()
            t, dto = ref.advance(t, 100.0, stepper, literal=True)
            assert dt == pytest.approx(dto, rel=1e-11), i
            for k in ("vortmax", "vortrms", "vorch", "ggmax", "umax", "vmax", "wmax", "usggmax", "lsggmax"):
                assert diag[k] == pytest.approx(ref.diag[k], rel=1e-10), (i, k)
            if i % 10 == 9 or i < 3:
                assert rel(lib.download3("svor"), ref.svor) < FIELD_TOL, i
        assert s.t == pytest.approx(t, rel=1e-11)
        lib.vor2vel()
        ref.vor2vel()
        for name in ("vor", "vel"):     # accumulated over nsteps steps, not a per-step figure
            assert rel(lib.download3(name), getattr(ref, name)) < 1e-10, name
        d = lib.diagnostics()
        assert d["ke"] == pytest.approx(ref.get_kinetic_energy(), rel=DIAG_TOL)
        assert d["en"] == pytest.approx(ref.get_enstrophy(), rel=DIAG_TOL)
        assert d["helicity"] == pytest.approx(ref.get_helicity(), rel=DIAG_TOL)
    finally:
        s.close()


@pytest.mark.parametrize("stepper", ["cn2", "impl-diff-rk4"])
def test_per_step_error_resynchronised(lib, stepper):
    """north_star bar proper: ONE step from identical states agrees to 1e-12 (max-norm relative).
    The device state is re-synchronised with the oracle's before every step so that the figure is
    a per-step error, not an accumulated one."""
    from ps3d_b200 import host
    n = 32
    ref = O.beltrami_setup(n)
    s = host.beltrami_solver(lib, n, stepper=stepper)
    try:
        # perturb so that many modes are active
        rng = np.random.default_rng(11)
        ref.svor += 1e-3 * rng.uniform(-1, 1, ref.svor.shape) * ref.filt[None]
        t = 0.0
        worst = 0.0
        for i in range(5):
            for c in range(3):
                lib.upload("svor", c, ref.svor[c])
            s.t = t
            dt, diag = s.advance()
            t, dto = ref.advance(t, 100.0, stepper, literal=True)
            assert dt == pytest.approx(dto, rel=1e-12)
            worst = max(worst, rel(lib.download3("svor"), ref.svor))
            if stepper == "cn2":
                # the source is a small difference of O(1) flux terms (near-Beltrami flow): its own max-norm
                # is ~1e2 below that of its terms, so its relative round-off is correspondingly larger.
                # (After an rk4 step svorts holds exp(+2 D dt)-scaled scratch values up to 1e23 at the
                # highest wavenumbers, where the literal combine/decompose of the oracle is itself only
                # good to 1e-7: not a meaningful comparison.)
                assert rel(lib.download3("svorts"), ref.svorts) < 1e-10
        assert worst < FIELD_TOL
    finally:
        s.close()


@pytest.mark.parametrize("stepper", ["cn2", "impl-diff-rk4"])
def test_separate_entry_points_equal_advance(lib, stepper):
    """vor2vel / adapt / source / step called one by one (the reference's module procedures, advance.f90:85-102)
    give the same step as ps3d_cuda_advance."""
    from ps3d_b200 import host
    ref = O.beltrami_setup(32)
    s = host.beltrami_solver(lib, 32, stepper=stepper)
    try:
        lib.vor2vel()
        dt, diag = lib.adapt(0.0, 100.0)
        lib.source()
        t = lib.step(0.0, dt)
        to, dto = ref.advance(0.0, 100.0, stepper, literal=True)
        assert dt == pytest.approx(dto, rel=1e-12) and t == pytest.approx(to, rel=1e-12)
        assert rel(lib.download3("svor"), ref.svor) < FIELD_TOL
    finally:
        s.close()


def test_full_size_properties_256(lib):
    """Properties that need no oracle at 256^3: inverse(forward) = identity, DST/DCT are
    self-inverse, linearity, combine/decompose are mutual inverses, Parseval for the packed FFT."""
    n = 256
    lib.init(n, n, n, -0.5 * PI * np.ones(3), PI * np.ones(3))
    lib.init_inversion("Hou & Li")
    try:
        rng = np.random.default_rng(3)
        f = rng.uniform(-1, 1, (n, n, n + 1))
        g = rng.uniform(-1, 1, (n, n, n + 1))
        fs = lib.fftxyp2s(f)
        assert rel(lib.fftxys2p(fs), f) < FIELD_TOL
        gs = lib.fftxyp2s(g)
        assert rel(lib.fftxyp2s(2.0 * f - 3.0 * g), 2.0 * fs - 3.0 * gs) < FIELD_TOL
        # Parseval: sum f^2 = sum_k w_k |packed|^2 with weight 2 except DC/Nyquist rows, per axis
        w = np.full(n, 2.0); w[0] = 1.0; w[n // 2] = 1.0
        lhs = np.sum(f[..., 5] ** 2)
        rhs = np.sum(w[:, None] * w[None, :] * fs[..., 5] ** 2)
        assert lhs == pytest.approx(rhs, rel=1e-12)
        d = lib.fftsine(f)
        dd = lib.fftsine(d)
        assert rel(dd[..., 1:n], f[..., 1:n]) < FIELD_TOL and np.all(dd[..., n] == 0.0)
        assert rel(lib.fftcosine(lib.fftcosine(f)), f) < FIELD_TOL
        assert rel(lib.field_decompose_semi_spectral(lib.field_combine_semi_spectral(f)), f) < FIELD_TOL
    finally:
        lib.finalise()


@pytest.mark.parametrize("stepper", ["cn2", "impl-diff-rk4"])
def test_config3_256_steps_match_oracle(lib, stepper):
    """BASELINE config 3 (Beltrami 256^3, full time stepping) against the oracle (literal steppers): dt to 1e-11 and
    the spectral vorticity to 1e-12 of its max-norm after each of the first steps."""
    from ps3d_b200 import host
    n, nsteps = 256, 2
    ref = O.beltrami_setup(n)
    s = host.beltrami_solver(lib, n, stepper=stepper)
    try:
        t = 0.0
        for i in range(nsteps):
            dt, diag = s.advance()
            t, dto = ref.advance(t, 100.0, stepper, literal=True)
            assert dt == pytest.approx(dto, rel=1e-11), i
            for k in ("vortmax", "vortrms", "vorch", "ggmax", "umax", "vmax", "wmax"):
                assert diag[k] == pytest.approx(ref.diag[k], rel=1e-10), (i, k)
            assert rel(lib.download3("svor"), ref.svor) < FIELD_TOL, i
        d = lib.diagnostics()         # of the vel / vor of the last vor2vel inside the step
        lib.vor2vel()
        ref.vor2vel()
        d = lib.diagnostics()
        assert d["ke"] == pytest.approx(ref.get_kinetic_energy(), rel=DIAG_TOL)
        assert d["en"] == pytest.approx(ref.get_enstrophy(), rel=DIAG_TOL)
        assert d["helicity"] == pytest.approx(ref.get_helicity(), rel=DIAG_TOL)
    finally:
        s.close()


def test_config4_512_analytic_known_answers(lib):
    """BASELINE config 4 grid (512^3) through the C ABI against analytic answers: unit-tests/test_vor2vel_1.f90
    (Beltrami k = l = 2, m = 1: velocity = vorticity / alpha) and test_diffx.f90 / test_diffy.f90
    (cos 4x -> -4 sin 4x, 1e-12), the reference's own known-answer tests at the benchmark size."""
    n = 512
    lower, extent = -0.5 * PI * np.ones(3), PI * np.ones(3)
    lib.init(n, n, n, lower, extent)
    lib.init_inversion("Hou & Li")
    try:
        from ps3d_b200 import host
        vor = host.beltrami_vorticity(n, n, n, lower, extent)
        alpha = 3.0
        lib.upload_vorticity(vor)
        lib.vor2vel()
        vor /= alpha                                   # analytic velocity (beltrami.f90:162-181, test_vor2vel_1.f90:75-94)
        err = 0.0
        for c in range(3):
            err = max(err, float(np.max(np.abs(lib.download("vel", c) - vor[c]))))
        del vor
        assert err < 1e-13, err                        # reference bar at 32^3: 1e-14 (test_vor2vel_1.f90:99)
        # diffx / diffy on a box of [-pi, pi]^3 need a second init
    finally:
        lib.finalise()
    lib.init(n, n, n, -PI * np.ones(3), 2 * PI * np.ones(3))
    lib.init_inversion("Hou & Li")
    try:
        x = (-PI + (2 * PI / n) * np.arange(n))
        f = np.empty((n, n, n + 1))
        f[:] = np.cos(4 * x)[:, None, None]
        out = lib.fftxys2p(lib.diffx(lib.fftxyp2s(f)))
        f[:] = (-4 * np.sin(4 * x))[:, None, None]
        assert np.max(np.abs(out - f)) < 1e-12
        f[:] = np.cos(4 * x)[None, :, None]
        out = lib.fftxys2p(lib.diffy(lib.fftxyp2s(f)))
        f[:] = (-4 * np.sin(4 * x))[None, :, None]
        assert np.max(np.abs(out - f)) < 1e-12
    finally:
        lib.finalise()


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_multi_gpu_slab_matches_oracle(nranks):
    """P-GPU slab decomposition with NCCL all-to-alls against the one-rank oracle (needs P GPUs)."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}", "--master-addr",
           "127.0.0.1", "--master-port", "29621", os.path.join(root, "tests", "multirank_worker.py"), "cuda", "cn2", "64"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("worst") == nranks


def test_multi_gpu_buoyancy_matches_oracle():
    """The ENABLE_BUOYANCY paths on a 2-GPU slab decomposition (peer-memory transport) against the one-rank oracle."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29623", os.path.join(root, "tests", "multirank_worker.py"), "cuda", "cn2", "32",
           "buoyancy"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("worst") == 2


def test_cpp_driver_matches_oracle(lib):
    """examples/ps3d_driver.cpp (C++ stand-in for ps3d.f90) through the same C ABI: 5 cn2 steps of Beltrami 32^3."""
    import os
    import re
    import subprocess
    import __graft_entry__ as G
    exe = G.build_driver()
    out = subprocess.run([exe, "--n", "32", "--steps", "5"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    ref = O.beltrami_setup(32)
    t = 0.0
    dts = []
    for _ in range(5):
        t, dt = ref.advance(t, 100.0, "cn2", literal=True)
        dts.append(dt)
    got = [float(m) for m in re.findall(r"and time step\s+(\S+)", out.stdout)]
    assert len(got) == 5 and all(g == pytest.approx(r, rel=1e-12) for g, r in zip(got, dts))
    ref.vor2vel()
    ke = float(re.search(r"ke (\S+)", out.stdout).group(1))
    assert ke == pytest.approx(ref.get_kinetic_energy(), rel=1e-12)


def test_streamed_upload_matches_blocking_upload(lib):
    """ps3d_cuda_upload_vorticity_begin/_end (copy stream + staging fields) against the blocking upload, with an
    advance running between _begin and _end as in bench.py's end-to-end loop."""
    from ps3d_b200 import host
    import torch
    n = 64
    s = host.beltrami_solver(lib, n, stepper="cn2")
    try:
        lower = -0.5 * PI * np.ones(3)
        extent = PI * np.ones(3)
        v = torch.from_numpy(host.beltrami_vorticity(n, n, n, lower, extent)).pin_memory().numpy()
        lib.upload_vorticity(v)
        s0 = lib.download3("svor")
        lib.upload_vorticity_begin(v)          # copies in flight while the step runs
        s.t = 0.0
        dt1, _ = s.advance()
        assert rel(lib.download3("svor"), s0) > 1e-6          # the step changed the state
        lib.upload_vorticity_end()                             # ... and the streamed upload restores the initial one
        assert np.array_equal(lib.download3("svor"), s0)
        s.t = 0.0
        dt2, _ = s.advance()
        assert dt1 == dt2
    finally:
        s.close()
