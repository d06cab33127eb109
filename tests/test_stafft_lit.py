"""Pins the oracle's library transforms (oracle/ps3d_oracle.py: numpy/scipy FFTs with the reference's packing)
to a literal C restatement of the reference's own transform code (oracle/stafft_lit.c <- src/fft/stafft.f90:
initfft, factorisen, forfft, revfft, dct, dst, forrdx4/3/2, revrdx4/3/2, including the sequential
post-processing recurrences :466-471, :526-533), for every power-of-two length the BASELINE configurations use,
and measures the round-off floor of the reference's arithmetic itself (the 1e-12 parity bar is judged against
it).  The Fortran cannot be compiled in this image; this is the closest pin to "the reference run here"."""
import math

import numpy as np
import pytest

import __graft_entry__ as G
from oracle import ps3d_oracle as O

G.build_stafft_lit()
from oracle.stafft_lit import Stafft  # noqa: E402

SIZES = [8, 16, 32, 64, 128, 256, 512, 1024]
TOL = 2e-14          # |oracle - literal|, white noise in [-1, 1] (measured <= 7e-15 at n = 1024)


@pytest.mark.parametrize("n", SIZES)
def test_oracle_transforms_equal_literal_stafft(n):
    S = Stafft(n)
    # factorisen: 6 first, then 4, 2, 3, 5 (stafft.f90:128-187) -> powers of two are 4^a 2^b
    assert S.factors[1] == int(math.log2(n)) // 2 and S.factors[2] == int(math.log2(n)) % 2
    rng = np.random.default_rng(n)
    x = rng.uniform(-1, 1, (6, n))
    f = S.forfft(x)
    assert np.max(np.abs(f - O.forfft(x, 1))) < TOL
    assert np.max(np.abs(S.revfft(f) - O.revfft(f, 1))) < TOL
    assert np.max(np.abs(S.revfft(f) - x)) < TOL                       # revfft(forfft(x)) = x
    xs = rng.uniform(-1, 1, (6, n))
    assert np.max(np.abs(S.dst(xs) - O.dst(xs, n))) < TOL
    xc = rng.uniform(-1, 1, (6, n + 1))
    assert np.max(np.abs(S.dct(xc) - O.dct(xc, n))) < TOL


def test_vector_call_equals_line_calls():
    """forfft(m, n, x(m, n)) with m > 1 (vector index fastest) == m calls with m = 1."""
    n, m = 64, 5
    S = Stafft(n)
    x = np.random.default_rng(1).uniform(-1, 1, (m, n))
    assert np.array_equal(S.forfft_m(x.T.copy()).T, S.forfft(x))


def test_radix3_lengths():
    """A length with factors 4, 2 and 3 but no 6 does not exist for even n (6 is tried first); 9 and 27 exercise
    forrdx3 / revrdx3 alone."""
    for n in (9, 27):
        S = Stafft(n)
        x = np.random.default_rng(n).uniform(-1, 1, (3, n))
        f = S.forfft(x)
        X = np.fft.rfft(x, axis=1) / math.sqrt(n)
        assert np.max(np.abs(f[:, : n // 2 + 1] - X.real)) < TOL
        assert np.max(np.abs(f[:, n // 2 + 1:] - X.imag[:, (n - 1) // 2:0:-1])) < TOL
        assert np.max(np.abs(S.revfft(f) - x)) < TOL


def test_reference_roundoff_floor():
    """The reference's own arithmetic: dct(dct(x)) - x and dst(dst(x)) - x grow with n because of the sequential
    recurrence x(2j+1) = x(2j-1) -+ rt2*wk(.) (stafft.f90:466-471, 526-533).  Recorded in DESIGN.md section 2."""
    floor = {}
    for n in SIZES:
        S = Stafft(n)
        rng = np.random.default_rng(7)
        xc = rng.uniform(-1, 1, (16, n + 1))
        xs = rng.uniform(-1, 1, (16, n))
        xs[:, n - 1] = 0.0
        floor[n] = (float(np.max(np.abs(S.dct(S.dct(xc)) - xc))), float(np.max(np.abs(S.dst(S.dst(xs)) - xs))))
    print("stafft round-off floor (dct.dct - I, dst.dst - I):", floor)
    assert floor[512][0] < 2e-13 and floor[1024][0] < 4e-13
    # the library transforms of the oracle are tighter than the reference's own arithmetic
    xc = np.random.default_rng(7).uniform(-1, 1, (16, 513))
    assert np.max(np.abs(O.dct(O.dct(xc, 512), 512) - xc)) < floor[512][0]


def test_fftxyp2s_equals_literal_2d():
    """fftxyp2s (sta3dfft.f90:136-194): forfft along y then along x of every z level, literal 1-D transforms."""
    nx, ny, nz = 32, 64, 4
    s = O.PS3D(nx, ny, nz, [0.0, 0.0, 0.0], [1.0, 1.0, 1.0])
    f = np.random.default_rng(2).uniform(-1, 1, (nx, ny, nz + 1))
    Sy, Sx = Stafft(ny), Stafft(nx)
    g = np.moveaxis(Sy.forfft(np.moveaxis(f, 1, -1)), -1, 1)       # y lines
    g = np.moveaxis(Sx.forfft(np.moveaxis(g, 0, -1)), -1, 0)       # x lines
    assert np.max(np.abs(g - s.fftxyp2s(f))) < TOL
    h = np.moveaxis(Sx.revfft(np.moveaxis(g, 0, -1)), -1, 0)
    h = np.moveaxis(Sy.revfft(np.moveaxis(h, 1, -1)), -1, 1)
    assert np.max(np.abs(h - f)) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(512, 8, 16), (8, 512, 16), (16, 16, 512), (256, 256, 8)])
def test_cuda_transforms_equal_literal_stafft(shape):
    """The CUDA transforms at the benchmarked line lengths against the literal stafft arithmetic (not the
    library FFTs): fftxyp2s, fftxys2p, fftsine, fftcosine through the C ABI."""
    import ps3d_b200
    lib = ps3d_b200.load()
    nx, ny, nz = shape
    lib.init(nx, ny, nz, np.zeros(3), np.ones(3))
    try:
        lib.init_inversion("Hou & Li")
        f = np.random.default_rng(5).uniform(-1, 1, (nx, ny, nz + 1))
        Sx, Sy, Sz = Stafft(nx), Stafft(ny), Stafft(nz)
        g = np.moveaxis(Sy.forfft(np.moveaxis(f, 1, -1)), -1, 1)
        g = np.moveaxis(Sx.forfft(np.moveaxis(g, 0, -1)), -1, 0)
        fs = lib.fftxyp2s(f)
        assert np.max(np.abs(fs - g)) < 1e-13
        assert np.max(np.abs(lib.fftxys2p(g) - f)) < 1e-13
        want = f.copy()
        want[..., 1:] = Sz.dst(f[..., 1:])                          # fftsine (sta3dfft.f90:264-276): rows 1..nz
        got = lib.fftsine(f)
        assert np.max(np.abs(got[..., 1:] - want[..., 1:])) < 1e-13
        assert np.max(np.abs(lib.fftcosine(f) - Sz.dct(f))) < 1e-13
    finally:
        lib.finalise()
