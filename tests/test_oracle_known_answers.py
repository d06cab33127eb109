"""Pin the oracle against the reference's own analytic known-answer tests
(/root/reference/unit-tests/*.f90).  The tolerances
are the reference's (file:line cited per test)."""
import math

import numpy as np
import pytest

from oracle import ps3d_oracle as O

PI = math.pi


def grid(s):
    x = (s.lower[0] + s.dx[0] * np.arange(s.nx))[:, None, None]
    y = (s.lower[1] + s.dx[1] * np.arange(s.ny))[None, :, None]
    z = (s.lower[2] + s.dx[2] * np.arange(s.nz + 1))[None, None, :]
    return x, y, z


def decompose_all(s, vor):
    s.vor[:] = vor
    for nc in range(3):
        s.svor[nc] = s.field_decompose_physical(vor[nc])


def test_forfft_revfft_roundtrip_and_packing():
    rng = np.random.default_rng(0)
    for n in (4, 8, 32, 64, 512):
        x = rng.uniform(-1, 1, (n, 3))
        y = O.forfft(x, 0)
        X = np.fft.fft(x, axis=0) / math.sqrt(n)
        assert np.allclose(y[0], X[0].real, atol=1e-15)
        assert np.allclose(y[n // 2], X[n // 2].real, atol=1e-15)
        for k in range(1, n // 2):
            assert np.allclose(y[k], X[k].real, atol=1e-14)
            assert np.allclose(y[n - k], X[k].imag, atol=1e-14)
        assert np.max(np.abs(O.revfft(y, 0) - x)) < 1e-14
    # stafft.f90:917-922 hand check: n=4 slot 3 holds (x3 - x1)/2
    x = np.array([1.0, 2.0, 5.0, 11.0])
    assert abs(O.forfft(x, 0)[3] - (x[3] - x[1]) / 2.0) < 1e-15


def test_dst_dct_definitions_and_self_inverse():
    rng = np.random.default_rng(1)
    for n in (8, 32, 512):
        x = rng.uniform(-1, 1, n + 1)
        j = np.arange(1, n)
        # stafft.f90:489-550
        S = np.array([math.sqrt(2.0 / n) * np.sum(x[1:n] * np.sin(PI * ((j * k) % (2 * n)) / n)) for k in range(1, n)])
        d = O.dst(x[1:], n)
        assert np.max(np.abs(d[: n - 1] - S)) < 1e-13 and d[n - 1] == 0.0
        assert np.max(np.abs(O.dst(d, n)[: n - 1] - x[1:n])) < 1e-13
        # stafft.f90:410-483
        C = np.array([math.sqrt(2.0 / n) * (0.5 * x[0] + np.sum(x[1:n] * np.cos(PI * ((j * k) % (2 * n)) / n))
                                            + 0.5 * (-1) ** k * x[n]) for k in range(n + 1)])
        c = O.dct(x, n)
        assert np.max(np.abs(c - C)) < 1e-13
        assert np.max(np.abs(O.dct(c, n) - x)) < 1e-13


def test_vor2vel_1_beltrami():
    """unit-tests/test_vor2vel_1.f90:99 (atol 1e-14)."""
    s = O.PS3D(32, 32, 32, -0.5 * PI * np.ones(3), PI * np.ones(3))
    x, y, z = grid(s)
    k, l, m = 2.0, 2.0, 1.0
    alpha = math.sqrt(k * k + l * l + m * m)
    f = 1.0 / (k * k + l * l)
    ref = np.empty((3, 32, 32, 33))
    ref[0] = f * (k * m * np.sin(m * z) - l * alpha * np.cos(m * z)) * np.sin(k * x + l * y)
    ref[1] = f * (l * m * np.sin(m * z) + k * alpha * np.cos(m * z)) * np.sin(k * x + l * y)
    ref[2] = np.cos(m * z) * np.cos(k * x + l * y)
    decompose_all(s, alpha * ref)
    s.vor2vel()
    assert np.max(np.abs(ref - s.vel)) < 1e-14
    # beltrami.f90:162-181 gives the same vorticity
    assert np.max(np.abs(O.beltrami_vorticity(32, 32, 32, s.lower, s.extent) - alpha * ref)) < 1e-14


def test_vor2vel_2_polynomial():
    """unit-tests/test_vor2vel_2.f90:104 (atol 1.2e-2)."""
    s = O.PS3D(32, 32, 32, [-0.5, -0.5, 0.0], [1.0, 1.0, 1.0])
    x, y, z = grid(s)
    l = 2 * PI
    k = 2 * l
    klsq = k * k - l * l
    f = 2 * z - z ** 2 - z ** 3
    dfdz = 2 - 2 * z - 3 * z ** 2
    d2 = -2 - 6 * z
    ref = np.empty((3, 32, 32, 33))
    ref[0] = k * dfdz * np.cos(k * x) * np.sin(l * y)
    ref[1] = -l * dfdz * np.sin(k * x) * np.cos(l * y)
    ref[2] = klsq * f * np.sin(k * x) * np.sin(l * y)
    vor = np.empty_like(ref)
    vor[0] = l * (klsq * f + d2) * np.sin(k * x) * np.cos(l * y)
    vor[1] = k * (d2 - klsq * f) * np.cos(k * x) * np.sin(l * y)
    vor[2] = -2 * k * l * dfdz * np.cos(k * x) * np.cos(l * y)
    decompose_all(s, vor)
    s.vor2vel()
    assert np.max(np.abs(ref - s.vel)) < 1.2e-2


@pytest.mark.parametrize("case,atol", [(3, 1e-15), (4, 1e-15), (5, 4e-6)])
def test_vor2vel_345_mean_flow(case, atol):
    """unit-tests/test_vor2vel_3.f90:82, _4.f90, _5.f90: eta = 1, 2z, 3z^2."""
    s = O.PS3D(32, 32, 32, [-0.5, -0.5, 0.0], [1.0, 1.0, 1.0])
    x, y, z = grid(s)
    zc = 0.5 * (s.lower[2] + s.upper[2])
    vor = np.zeros((3, 32, 32, 33))
    ref = np.zeros_like(vor)
    if case == 3:
        vor[1] = 1.0 + 0 * z
        ref[0] = z - zc + 0 * x * y
    elif case == 4:
        vor[1] = 2 * z + 0 * x * y
        ref[0] = z ** 2 - 1.0 / 3.0 + 0 * x * y
    else:
        vor[1] = 3 * z ** 2 + 0 * x * y
        ref[0] = z ** 3 - 0.25 + 0 * x * y
    decompose_all(s, vor)
    s.vor2vel()
    assert np.max(np.abs(ref - s.vel)) < max(atol, 2e-15)


def test_diffx_diffy():
    """unit-tests/test_diffx.f90:75, test_diffy.f90 (1e-12)."""
    s = O.PS3D(64, 128, 64, -PI * np.ones(3), 2 * PI * np.ones(3))
    x, y, z = grid(s)
    fp = np.cos(4 * x) + 0 * y + 0 * z
    out = s.fftxys2p(s.diffx(s.fftxyp2s(fp)))
    assert np.max(np.abs(out + 4 * np.sin(4 * x))) < 1e-12
    s = O.PS3D(128, 64, 64, -PI * np.ones(3), 2 * PI * np.ones(3))
    x, y, z = grid(s)
    fp = np.cos(4 * y) + 0 * x + 0 * z
    out = s.fftxys2p(s.diffy(s.fftxyp2s(fp)))
    assert np.max(np.abs(out + 4 * np.sin(4 * y))) < 1e-12


def test_diffz():
    """unit-tests/test_diffz_1.f90 (z -> 1, 1e-14), test_diffz_2.f90:78 (2e-14)."""
    s = O.PS3D(32, 32, 32, [0.0, 0.0, 0.0], [1.0, 1.0, 1.0])
    x, y, z = grid(s)
    assert np.max(np.abs(s.central_diffz(z + 0 * x * y) - 1.0)) < 1e-14
    s = O.PS3D(16, 32, 32, [0.0, -0.5 * PI, 0.0], [PI, 2 * PI, 2 * PI])
    x, y, z = grid(s)
    f = z * np.cos(2 * x) * np.sin(y)
    assert np.max(np.abs(s.central_diffz(f) - np.cos(2 * x) * np.sin(y))) < 2e-14


def test_diffz_3_4():
    """unit-tests/test_diffz_3.f90:52-63 (6z + 3 cos(kx) sin(ly) z, atol 2e-13) and test_diffz_4.f90 (1 - 3z^2 ->
    -6z with the one-sided boundary rows, atol 1e-1)."""
    s = O.PS3D(16, 32, 32, [0.0, -0.5 * PI, 0.0], [PI, 2 * PI, 2 * PI])
    x, y, z = grid(s)
    f = 6 * z + 3 * np.cos(2 * x) * np.sin(y) * z
    assert np.max(np.abs(s.central_diffz(f) - (6 + 3 * np.cos(2 * x) * np.sin(y)))) < 2e-13
    s = O.PS3D(32, 32, 32, [0.0, 0.0, 0.0], [1.0, 1.0, 1.0])
    x, y, z = grid(s)
    assert np.max(np.abs(s.central_diffz(1 - 3 * z ** 2 + 0 * x * y) + 6 * z)) < 1e-1


def deriv_via_sine_series(nz, dst_fn, dct_fn):
    """tests/test_deriv.f90:33-62: d/dz of phi = 1 - 2 z^3 on [0, 1] from the sine series of phi - phi_lin
    (dst -> * rkz -> dct, + slope of phi_lin); returns (spectral derivative, exact derivative)."""
    z = np.arange(nz + 1) / nz
    a = (1 - 2 * z ** 3) - (1 - 2 * z)
    w = dst_fn(a)                                   # rows 1..nz-1 transformed, row nz -> 0
    rkz = PI * np.arange(nz + 1)                    # deriv1d.f90:17-22 with Lz = 1
    d = dct_fn(rkz * w) - 2.0
    return d, -6 * z ** 2


@pytest.mark.parametrize("nz", [64, 1024])
def test_deriv(nz):
    """tests/test_deriv.f90 (prints the errors, no tolerance): the decomposition-based derivative converges like
    3/nz in the max norm (2.93e-3 at the reference's nz = 1024)."""
    def dst_fn(a):
        out = np.zeros(nz + 1)
        out[1:] = O.dst(a[1:].copy(), nz)
        return out
    d, exact = deriv_via_sine_series(nz, dst_fn, lambda v: O.dct(v, nz))
    assert np.max(np.abs(d - exact)) < 3.01 / nz


def test_implicit_rk():
    """unit-tests/test_implicit_rk.f90:64-94: S = D = 1, exact solution
    q E + (1 - E) S / D with E = exp(-dt D); atol 1e-12 at dt = 0.0125."""
    s = O.PS3D(32, 32, 32, [0.0, 0.0, 0.0], [1.0, 1.0, 1.0])
    s.nnu = 3
    s.vor2vel = lambda: None
    def reset_source():          # test_implicit_rk.f90:150 `svorts = src`
        s.svorts[:] = 1.0
    s.source = reset_source
    s.adjust_vorticity_mean = lambda: None
    s.filt[:] = 1.0
    dt = 0.1
    for _ in range(4):
        s.svor[:] = 2.0
        s.svorts[:] = 1.0
        ed = math.exp(-dt)
        ref = ed * 2.0 + (1.0 - ed) * 1.0
        # impl_rk4_set_diffusion: vdiss = dt/2 * vorch * vhdis with D = 1
        s.vdiss = 0.5 * dt * np.ones((32, 32))
        s.rk4_step(0.0, dt, literal=True)
        err = np.max(np.abs(s.svor - ref))
        dt *= 0.5
    assert err < 1e-12


def test_combine_decompose_identity():
    """decompose(c(ky,kx) * combine(q)) == c*q to round-off."""
    s = O.PS3D(32, 32, 32, -0.5 * PI * np.ones(3), PI * np.ones(3))
    rng = np.random.default_rng(5)
    q = rng.uniform(-1, 1, (32, 32, 33))
    c = rng.uniform(0.1, 1, (32, 32))
    out = s.field_decompose_semi_spectral(c[..., None] * s.field_combine_semi_spectral(q))
    assert np.max(np.abs(out - c[..., None] * q)) < 2e-14


def test_jacobi_vs_lapack():
    rng = np.random.default_rng(7)
    a = rng.uniform(-1, 1, (6, 500))
    d = O.jacobi_eigenvalues(*a)
    lam = np.maximum.reduce([np.abs(v) for v in d])
    M = np.zeros((500, 3, 3))
    M[:, 0, 0], M[:, 0, 1], M[:, 0, 2], M[:, 1, 1], M[:, 1, 2], M[:, 2, 2] = a
    M[:, 1, 0], M[:, 2, 0], M[:, 2, 1] = a[1], a[2], a[4]
    ref = np.max(np.abs(np.linalg.eigvalsh(M)), axis=1)
    assert np.max(np.abs(lam - ref)) < 5e-15


def test_literal_and_fused_steppers_agree():
    """The literal steppers (combine -> x c(ky,kx) -> decompose pairs as in
    cn2.f90:124-135 / impl_rk4.f90:232-238) and the fused ones (identity a11)
    agree to round-off over a few steps.  (The discrete source is not ~0 for
    Beltrami: vor2vel's solenoidal projection uses the 2nd-order
    central_diffz, so vor differs from alpha*vel by O(dz^2).)"""
    a = O.beltrami_setup(32)
    b = O.beltrami_setup(32)
    ta = tb = 0.0
    for _ in range(3):
        ta, _ = a.advance(ta, 100.0, "cn2", literal=True)
        tb, _ = b.advance(tb, 100.0, "cn2", literal=False)
    assert ta == pytest.approx(tb, rel=1e-13)
    assert np.max(np.abs(a.svor - b.svor)) < 1e-13 * np.max(np.abs(a.svor))
    a = O.beltrami_setup(32)
    b = O.beltrami_setup(32)
    ta = tb = 0.0
    for _ in range(2):
        ta, _ = a.advance(ta, 100.0, "impl-diff-rk4", literal=True)
        tb, _ = b.advance(tb, 100.0, "impl-diff-rk4", literal=False)
    assert np.max(np.abs(a.svor - b.svor)) < 1e-13 * np.max(np.abs(a.svor))
