"""CPU oracle for the ps3d time-step hot path.  TEST INFRASTRUCTURE ONLY.

This file is a NumPy/SciPy *restatement* of the reference algorithm
(matt-frey/ps3d v0.1.3, Fortran).  It exists so that the CUDA path can be
checked against something that follows the reference line by line.  Only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import it.  The product (`ps3d_b200`) never does.

Parity pinning: the reference cannot be compiled here (no gfortran / MPI /
netCDF) and ships no golden vectors, so the oracle is pinned by (a) the
reference's own analytic known-answer unit tests (tests/test_oracle_*.py):
test_vor2vel_1..5, test_diffx/diffy, test_diffz_1..4, test_implicit_rk,
DST/DCT self-inverse and revfft(forfft(x)) = x, and (b) a literal C
restatement of the reference's own transform code, oracle/stafft_lit.c <-
src/fft/stafft.f90 (tests/test_stafft_lit.py: every power-of-two length
8..1024, <= 2e-14): pinned at operator level.  Multi-step trajectories are
pinned by nothing but this oracle and its independent C++ twin
(oracle/ps3d_ref.cpp; the reference has no such test): trajectory parity is
"oracle-pinned, reference-unpinned".  The ENABLE_BUOYANCY paths have no test
in the reference at all; they are pinned by an analytic known answer for
diffz (tests/test_buoyancy.py) and by inspection against the cited lines.

Array convention: Fortran `f(0:nz, y, x)` == NumPy `f[x, y, z]` (C order, z
contiguous), `f(0:nz, y, x, c)` == `f[c, x, y, z]`.

Transforms use library FFTs with the reference's packing/normalisation
(verified against a transliteration of stafft.f90 to <= 2e-14, SURVEY.md
section 8c); this deliberately does NOT reproduce the round-off of the
sequential post-processing recurrence in stafft.f90:466-471.
"""
from __future__ import annotations

import math
import numpy as np
import scipy.fft as sfft

_WORKERS = -1          # scipy.fft worker threads (all cores)
SMALL = 1.0e-12        # constants.f90:65
CFLMAX = 0.8           # constants.f90:63


# --------------------------------------------------------------------------
# 1-D transforms (stafft.f90)
# --------------------------------------------------------------------------
def forfft(x: np.ndarray, axis: int) -> np.ndarray:
    """stafft.f90:196-287 `forfft`: real FFT, Hermitian-packed, 1/sqrt(n).

    y[k] = Re X_k/sqrt(n) (k=0..n/2), y[n-k] = Im X_k/sqrt(n) (k=1..n/2-1),
    X_k = sum_j x_j exp(-2 pi i j k / n).
    """
    x = np.moveaxis(np.asarray(x, dtype=np.float64), axis, 0)
    n = x.shape[0]
    X = sfft.rfft(x, axis=0, workers=_WORKERS) / math.sqrt(n)
    y = np.empty_like(x)
    y[: n // 2 + 1] = X.real
    if n > 2:
        y[n // 2 + 1:] = X.imag[n // 2 - 1:0:-1]
    return np.moveaxis(y, 0, axis)


def revfft(y: np.ndarray, axis: int) -> np.ndarray:
    """stafft.f90:296-403 `revfft`: exact inverse of `forfft`."""
    y = np.moveaxis(np.asarray(y, dtype=np.float64), axis, 0)
    n = y.shape[0]
    X = np.zeros((n // 2 + 1,) + y.shape[1:], dtype=np.complex128)
    X.real[:] = y[: n // 2 + 1]
    if n > 2:
        X.imag[1: n // 2] = y[n - 1: n // 2: -1]
    x = sfft.irfft(X * math.sqrt(n), n=n, axis=0, workers=_WORKERS)
    return np.moveaxis(x, 0, axis)


def dst(x: np.ndarray, n: int) -> np.ndarray:
    """stafft.f90:489-550 `dst(1, n, x(1:n))` on the last axis.

    `x` holds slots 1..n (length n).  DST-I of slots 1..n-1 scaled
    sqrt(2/n); slot n is never read (`:509-513`) and is set to 0 (`:546-549`).
    """
    out = np.zeros_like(x)
    out[..., : n - 1] = sfft.dst(x[..., : n - 1], type=1, axis=-1,
                                 workers=_WORKERS) / math.sqrt(2.0 * n)
    return out


def dct(x: np.ndarray, n: int) -> np.ndarray:
    """stafft.f90:410-483 `dct(1, n, x(0:n))` on the last axis (DCT-I,
    scaled sqrt(2/n), self-inverse)."""
    return sfft.dct(x, type=1, axis=-1, workers=_WORKERS) / math.sqrt(2.0 * n)


# --------------------------------------------------------------------------
# Jacobi eigenvalues of symmetric 3x3 (jacobi.f90)
# --------------------------------------------------------------------------
def _givens(aij, di, dj):
    """jacobi.f90:19-51 `givens` (vectorised)."""
    eps = np.finfo(np.float64).eps
    g = 100.0 * np.abs(aij)
    h = dj - di
    small_rot = (np.abs(h) + g) == np.abs(h)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        t1 = aij / (h + np.copysign(eps, h))
        theta = 0.5 * h / (aij + np.copysign(eps, aij))
        t2 = 1.0 / (np.abs(theta) + np.sqrt(1.0 + theta * theta))
        t2 = np.where(theta < 0.0, -t2, t2)
    t = np.where(small_rot, t1, t2)
    c = 1.0 / np.sqrt(1.0 + t * t)
    s = t * c
    tau = s / (1.0 + c)
    return c, s, t, tau


def jacobi_eigenvalues(s11, s12, s13, s22, s23, s33, atol=1.0e-15):
    """jacobi.f90:270-305 `jacobi_eigenvalues` for arrays of symmetric 3x3
    matrices (upper triangle given).  Returns (d1, d2, d3) unsorted — the
    caller only needs max|lambda| (advance.f90:264).

    The cyclic sweep order (1,2),(1,3),(2,3) and the Rutishauser update
    (`apply_rotation`, jacobi.f90:215-259) are followed literally; matrices
    whose off-diagonal sum is already <= atol are frozen (the Fortran `do
    while (sm > atol)` loop is per matrix).
    """
    a12 = np.array(s12, dtype=np.float64).ravel()
    a13 = np.array(s13, dtype=np.float64).ravel()
    a23 = np.array(s23, dtype=np.float64).ravel()
    d = [np.array(v, dtype=np.float64).ravel() for v in (s11, s22, s33)]
    b = [v.copy() for v in d]
    idx = np.arange(a12.size)
    sm = np.abs(a12) + np.abs(a13) + np.abs(a23)
    act = idx[sm > atol]
    sweeps = 0
    while act.size:
        A12, A13, A23 = a12[act], a13[act], a23[act]
        D = [d[0][act], d[1][act], d[2][act]]
        Z = [np.zeros_like(A12) for _ in range(3)]
        # (i, j) = (1, 2): k = 3 -> g = A(1,3), h = A(2,3)
        c, s, t, tau = _givens(A12, D[0], D[1])
        h = t * A12
        Z[0] -= h; Z[1] += h; D[0] = D[0] - h; D[1] = D[1] + h
        A12 = np.zeros_like(A12)
        g, hh = A13, A23
        A13 = g - s * (hh + g * tau)
        A23 = hh + s * (g - hh * tau)
        # (i, j) = (1, 3): k = 2 -> g = A(1,2), h = A(2,3)
        c, s, t, tau = _givens(A13, D[0], D[2])
        h = t * A13
        Z[0] -= h; Z[2] += h; D[0] = D[0] - h; D[2] = D[2] + h
        A13 = np.zeros_like(A13)
        g, hh = A12, A23
        A12 = g - s * (hh + g * tau)
        A23 = hh + s * (g - hh * tau)
        # (i, j) = (2, 3): k = 1 -> g = A(1,2), h = A(1,3)
        c, s, t, tau = _givens(A23, D[1], D[2])
        h = t * A23
        Z[1] -= h; Z[2] += h; D[1] = D[1] - h; D[2] = D[2] + h
        A23 = np.zeros_like(A23)
        g, hh = A12, A13
        A12 = g - s * (hh + g * tau)
        A13 = hh + s * (g - hh * tau)
        # B = B + Z; D = B; Z = 0
        for k in range(3):
            b[k][act] += Z[k]
            d[k][act] = b[k][act]
        a12[act], a13[act], a23[act] = A12, A13, A23
        sm = np.abs(A12) + np.abs(A13) + np.abs(A23)
        act = act[sm > atol]
        sweeps += 1
        if sweeps > 100:
            raise RuntimeError("Jacobi did not converge")
    return d[0], d[1], d[2]


class RollingMean:
    """rolling_mean.f90:36-69 `rolling_mean_t%get_next`."""

    def __init__(self, n: int):
        self.length = n
        self.history = np.zeros(n)
        self.inew = 1
        self.iold = 1
        self.sma = 0.0
        self.filled = False

    def get_next(self, vnew: float) -> float:
        if self.filled:
            vold = self.history[self.iold - 1]
            self.iold = self.iold % self.length + 1
            self.sma = self.sma + (vnew - vold) / float(self.length)
            self.history[self.inew - 1] = vnew
            self.inew = self.inew % self.length + 1
        else:
            self.history[self.inew - 1] = vnew
            self.sma = float(np.sum(self.history[: self.inew])) / float(self.inew)
            self.filled = self.length == self.inew
            self.inew = self.inew % self.length + 1
        return self.sma


# --------------------------------------------------------------------------
# The solver state + operators
# --------------------------------------------------------------------------
class PS3D:
    """Restatement of modules sta3dfft, inversion_utils, inversion_mod,
    field_diagnostics, advance_mod, cn2_mod, impl_rk4_mod on one rank."""

    def __init__(self, nx, ny, nz, lower, extent, filtering="Hou & Li"):
        self.nx, self.ny, self.nz = int(nx), int(ny), int(nz)
        self.lower = np.asarray(lower, dtype=np.float64)
        self.extent = np.asarray(extent, dtype=np.float64)
        # parameters.f90:61-86 update_parameters
        self.upper = self.lower + self.extent
        self.dx = self.extent / np.array([nx, ny, nz], dtype=np.float64)
        self.dxi = 1.0 / self.dx
        self.ncell = nx * ny * nz
        self.ncelli = 1.0 / float(self.ncell)
        self.fnzi = 1.0 / float(nz)
        self.filtering = filtering
        self._init_fft()
        self._init_inversion()
        shp = (3, nx, ny, nz + 1)
        # fields.f90:17-35
        self.svor = np.zeros(shp)
        self.vor = np.zeros(shp)
        self.vel = np.zeros(shp)
        self.svel = np.zeros(shp)
        self.svorts = np.zeros(shp)
        self.vortsm = np.zeros(shp)
        self.vdiss = np.zeros((nx, ny))
        self.ini_vor_mean = np.zeros(2)
        self.vhdis = None
        self.vvisc = None
        self.rollmean = None
        self.f_cor = np.zeros(3)
        # ENABLE_BUOYANCY build (configure.ac:228-245): buoyancy perturbation b' (fields.f90:28-35), squared buoyancy
        # frequency (physics.f90:92); off unless enable_buoyancy() is called
        self.buoyancy = False
        self.bfsq = 0.0
        self.buoy_rollmean = None
        self.diag = {}
        # impl_rk4 work arrays
        self.svori = None
        self.svorf = None

    # ---- sta3dfft.f90:53-110, sta2dfft.f90:44-90, deriv1d.f90:10-33 ----
    def _init_fft(self):
        nx, ny, nz = self.nx, self.ny, self.nz
        hrkx = math.pi / self.extent[0] * np.arange(1, nx + 1)   # hrkx(1:nx)
        hrky = math.pi / self.extent[1] * np.arange(1, ny + 1)
        self.hrkx, self.hrky = hrkx, hrky
        nwx, nwy = nx // 2, ny // 2
        rkx = np.zeros(nx)
        for k in range(1, nwx):
            rkx[k] = hrkx[2 * k - 1]
            rkx[nx - k] = hrkx[2 * k - 1]
        rkx[nwx] = hrkx[nx - 1]
        rky = np.zeros(ny)
        for k in range(1, nwy):
            rky[k] = hrky[2 * k - 1]
            rky[ny - k] = hrky[2 * k - 1]
        rky[nwy] = hrky[ny - 1]
        self.rkx, self.rky = rkx, rky
        rkz = np.zeros(nz + 1)
        rkz[1:] = math.pi / self.extent[2] * np.arange(1, nz + 1)
        self.rkz = rkz
        self.rkzi = 1.0 / rkz[1:nz]          # rkzi(1:nz-1)

    # ---- sta3dfft.f90:136-260 ----
    def fftxyp2s(self, fp):
        """`fftxyp2s`: forfft along y then along x (sta3dfft.f90:161-178)."""
        return forfft(forfft(fp, axis=1), axis=0)

    def fftxys2p(self, fs):
        """`fftxys2p`: revfft along x then along y (sta3dfft.f90:232-249)."""
        return revfft(revfft(fs, axis=0), axis=1)

    def fftsine(self, fs):
        """sta3dfft.f90:264-278: dst on slots 1..nz of every column."""
        out = fs.copy()
        out[..., 1:] = dst(fs[..., 1:], self.nz)
        return out

    def fftcosine(self, fs):
        """sta3dfft.f90:282-296."""
        return dct(fs, self.nz)

    # ---- sta3dfft.f90:304-377 ----
    def diffx(self, fs):
        """`diffx`: ds(kx) = si*hrkx(dkx)*fs(nx-kx); ds(0)=ds(nx/2)=0."""
        nx = self.nx
        nwx = nx // 2
        ds = np.zeros_like(fs)
        for kx in range(1, nx):
            dkx = min(2 * kx, 2 * (nx - kx))
            si = 1.0 if kx >= nwx + 1 else -1.0
            ds[kx] = si * self.hrkx[dkx - 1] * fs[nx - kx]
        if nx % 2 == 0:
            ds[nwx] = 0.0
        return ds

    def diffy(self, fs):
        """`diffy` (sta3dfft.f90:345-377)."""
        ny = self.ny
        nwy = ny // 2
        ds = np.zeros_like(fs)
        for ky in range(1, ny):
            dky = min(2 * ky, 2 * (ny - ky))
            si = 1.0 if ky >= nwy + 1 else -1.0
            ds[:, ky] = si * self.hrky[dky - 1] * fs[:, ny - ky]
        if ny % 2 == 0:
            ds[:, nwy] = 0.0
        return ds

    # ---- inversion_utils.f90:222-372 ----
    def _init_inversion(self):
        nx, ny, nz = self.nx, self.ny, self.nz
        self.dzi = self.dxi[2]
        self.hdzi = 0.5 * self.dxi[2]
        k2l2 = self.rkx[:, None] ** 2 + self.rky[None, :] ** 2      # [kx, ky]
        k2l2[0, 0] = 1.0
        k2l2i = 1.0 / k2l2
        k2l2[0, 0] = 0.0
        k2l2i[0, 0] = 0.0
        self.k2l2, self.k2l2i = k2l2, k2l2i
        if self.filtering == "2/3-rule":
            self._init_23rd_rule_filter()
        else:
            self._init_hou_and_li_filter()
        self.filt[0, 0, :] = 1.0
        green = np.empty((nx, ny, nz + 1))
        green[..., 1:] = -1.0 / (k2l2[..., None] + self.rkz[None, None, 1:] ** 2)
        green[..., 0] = -k2l2i
        self.green = green
        z = self.lower[2] + self.dx[2] * np.arange(nz + 1)
        zm = self.upper[2] - z
        zp = z - self.lower[2]
        self._set_hyperbolic_functions(zm, zp)
        self.phim[0, 0] = zm / self.extent[2]
        self.phip[0, 0] = zp / self.extent[2]
        self.dphim[0, 0] = -1.0 / self.extent[2]        # inversion_utils.f90:333-336 (ENABLE_BUOYANCY)
        self.dphip[0, 0] = 1.0 / self.extent[2]
        self.thetam[0, 0] = 0.0
        self.thetap[0, 0] = 0.0
        self.dthetam[0, 0] = 0.0
        self.dthetap[0, 0] = 0.0
        phip00 = self.phip[0, 0].copy()
        self.gamtop = 0.5 * self.extent[2] * (phip00 ** 2 - 1.0 / 3.0)
        self.gambot = self.gamtop[::-1].copy()

    def _init_hou_and_li_filter(self):
        """inversion_utils.f90:377-403."""
        nz = self.nz
        skx = -36.0 * (self.rkx / self.rkx.max()) ** 36
        sky = -36.0 * (self.rky / self.rky.max()) ** 36
        skz = -36.0 * (self.rkz / self.rkz.max()) ** 36
        f2 = np.exp(skx[:, None] + sky[None, :])
        filt = np.empty((self.nx, self.ny, nz + 1))
        filt[..., 0] = f2
        filt[..., nz] = f2
        filt[..., 1:nz] = f2[..., None] * np.exp(skz[1:nz])[None, None, :]
        self.filt = filt

    def _init_23rd_rule_filter(self):
        """inversion_utils.f90:408-455."""
        nz = self.nz
        f23 = 2.0 / 3.0
        skx = (self.rkx <= f23 * self.rkx.max()).astype(np.float64)
        sky = (self.rky <= f23 * self.rky.max()).astype(np.float64)
        skz = (self.rkz <= f23 * self.rkz.max()).astype(np.float64)
        f2 = skx[:, None] * sky[None, :]
        filt = np.empty((self.nx, self.ny, nz + 1))
        filt[..., 0] = f2
        filt[..., nz] = f2
        filt[..., 1:nz] = f2[..., None] * skz[None, None, 1:nz]
        self.filt = filt

    def _set_hyperbolic_functions(self, zm, zp):
        """inversion_utils.f90:484-542 (release build, no NDEBUG clamps);
        the (0,0) column is overwritten by the caller (`:326-346`)."""
        k2 = self.k2l2.copy()
        k2[0, 0] = 1.0        # placeholder, overwritten afterwards
        kl = np.sqrt(k2)[..., None]
        fac = kl * self.extent[2]
        ef = np.exp(-fac)
        div = 1.0 / (1.0 - ef ** 2)
        k2ifac = 0.5 * self.k2l2i[..., None]
        Lm = kl * zm[None, None, :]
        Lp = kl * zp[None, None, :]
        ep = np.exp(-Lp)
        em = np.exp(-Lm)
        self.phim = div * (ep - ef * em)
        self.phip = div * (em - ef * ep)
        dphim = -kl * div * (ep + ef * em)
        dphip = kl * div * (em + ef * ep)
        self.dphim, self.dphip = dphim, dphip          # module arrays in the ENABLE_BUOYANCY build (:520-521)
        Q = div * (1.0 + ef ** 2)
        R = div * 2.0 * ef
        self.thetam = k2ifac * (R * Lm * self.phip - Q * Lp * self.phim)
        self.thetap = k2ifac * (R * Lp * self.phim - Q * Lm * self.phip)
        self.dthetam = -k2ifac * ((Q * Lp - 1.0) * dphim - R * Lm * dphip)
        self.dthetap = -k2ifac * ((Q * Lm - 1.0) * dphip - R * Lp * dphim)

    # ---- inversion_utils.f90:124-218 ----
    def init_diffusion(self, te, en, nnu=3, prediss=30.0, length_scale="Kolmogorov"):
        K2max = max(self.rkx.max(), self.rky.max()) ** 2
        rkmsi = 1.0 / K2max
        if length_scale == "Kolmogorov":
            vis = prediss * (K2max * te / en) ** (1.0 / 3.0) * rkmsi ** nnu
        elif length_scale == "geophysical":
            vis = prediss * rkmsi ** nnu
        else:
            raise ValueError("We only support 'Kolmogorov' or 'geophysical'")
        self.vvisc = vis
        self.nnu = nnu
        if nnu == 1:
            self.vhdis = vis * self.k2l2
        else:
            self.vhdis = vis * self.k2l2 ** nnu
        return vis

    # ---- inversion_utils.f90:549-673 ----
    def field_decompose_semi_spectral(self, sfc):
        nz = self.nz
        out = sfc.copy()
        out[..., 1:nz] = sfc[..., 1:nz] - (sfc[..., 0:1] * self.phim[..., 1:nz]
                                           + sfc[..., nz:nz + 1] * self.phip[..., 1:nz])
        top = sfc[..., nz].copy()
        out[..., 1:] = dst(out[..., 1:], nz)
        out[..., nz] = top
        return out

    def field_combine_semi_spectral(self, sf):
        nz = self.nz
        out = sf.copy()
        top = sf[..., nz].copy()
        out[..., nz] = 0.0
        out[..., 1:] = dst(out[..., 1:], nz)
        out[..., nz] = top
        out[..., 1:nz] = (out[..., 1:nz] + out[..., 0:1] * self.phim[..., 1:nz]
                          + out[..., nz:nz + 1] * self.phip[..., 1:nz])
        return out

    def field_decompose_physical(self, fc):
        return self.field_decompose_semi_spectral(self.fftxyp2s(fc))

    def field_combine_physical(self, sf):
        return self.fftxys2p(self.field_combine_semi_spectral(sf))

    def central_diffz(self, fs):
        nz = self.nz
        ds = np.empty_like(fs)
        ds[..., 0] = self.dzi * (fs[..., 1] - fs[..., 0])
        ds[..., nz] = self.dzi * (fs[..., nz] - fs[..., nz - 1])
        ds[..., 1:nz] = (fs[..., 2:] - fs[..., : nz - 1]) * self.hdzi
        return ds

    # ---- inversion.f90:23-226 ----
    def vor2vel(self):
        nz = self.nz
        svor = self.svor
        k2l2i = self.k2l2i[..., None]
        as_ = self.diffx(svor[1])
        bs = self.diffy(svor[0])
        ds = as_ - bs
        cs = self.field_combine_semi_spectral(svor[2])
        es = self.central_diffz(cs)
        es = self.field_decompose_semi_spectral(es)
        ubar = svor[0, 0, 0, :].copy()
        vbar = svor[1, 0, 0, :].copy()
        svor[0] = k2l2i * (self.diffx(es) + self.diffy(ds))
        svor[1] = k2l2i * (self.diffy(es) - self.diffx(ds))
        svor[0, 0, 0, :] = ubar
        svor[1, 0, 0, :] = vbar
        for nc in range(3):
            self.vor[nc] = self.field_combine_physical(svor[nc])
        ds = self.diffy(svor[0]) - self.diffx(svor[1])
        bs = np.zeros_like(ds)
        bs[..., 1:nz] = (ds[..., 0:1] * self.thetam[..., 1:nz]
                         + ds[..., nz:nz + 1] * self.thetap[..., 1:nz])
        es = ds[..., 0:1] * self.dthetam + ds[..., nz:nz + 1] * self.dthetap
        ds[..., 1:nz] = self.green[..., 1:nz] * ds[..., 1:nz]
        as_ = np.zeros_like(ds)
        as_[..., 1:nz] = self.rkz[1:nz] * ds[..., 1:nz]
        as_ = dct(as_, nz)
        ds[..., 1:] = dst(ds[..., 1:], nz)
        ds[..., 0] = 0.0
        ds[..., 1:nz] = ds[..., 1:nz] + bs[..., 1:nz]
        ds[..., nz] = 0.0
        es = es + as_
        cs = self.field_combine_semi_spectral(svor[2])
        # horizontally averaged flow (inversion.f90:150-165)
        ubar = np.zeros(nz + 1)
        vbar = np.zeros(nz + 1)
        ubar[1:nz] = -self.rkzi * svor[1, 0, 0, 1:nz]
        vbar[1:nz] = self.rkzi * svor[0, 0, 0, 1:nz]
        ubar = dct(ubar, nz)
        vbar = dct(vbar, nz)
        ubar = ubar + svor[1, 0, 0, nz] * self.gamtop - svor[1, 0, 0, 0] * self.gambot
        vbar = vbar - svor[0, 0, 0, nz] * self.gamtop + svor[0, 0, 0, 0] * self.gambot
        as_ = k2l2i * (self.diffx(es) + self.diffy(cs))
        as_[0, 0, :] = ubar
        self.svel[0] = as_
        self.vel[0] = self.fftxys2p(as_)
        as_ = k2l2i * (self.diffy(es) - self.diffx(cs))
        as_[0, 0, :] = vbar
        self.svel[1] = as_
        self.vel[1] = self.fftxys2p(as_)
        self.svel[2] = ds
        self.vel[2] = self.fftxys2p(ds)

    # ---- inversion.f90:298-371 ----
    def vorticity_tendency(self):
        vel, vor = self.vel, self.vor
        for nc in range(3):
            vor[nc] = vor[nc] + self.f_cor[nc]
        if self.buoyancy:
            self.buoy = self.field_combine_physical(self.sbuoy)          # inversion.f90:316-318
        fp = vel[0] * vor[1] - vel[1] * vor[0]
        if self.buoyancy:
            fp = fp + self.buoy                                          # r = u * eta - v * xi + b (:329-331)
        r = self.field_decompose_physical(fp)
        fp = vel[2] * vor[0] - vel[0] * vor[2]
        q = self.field_decompose_physical(fp)
        s1 = self.diffy(r)
        p = self.field_decompose_physical(self.central_diffz(fp))
        self.svorts[0] = s1 - p
        fp = vel[1] * vor[2] - vel[2] * vor[1]
        p = self.field_decompose_physical(fp)
        s2 = self.diffx(r)
        r = self.field_decompose_physical(self.central_diffz(fp))
        self.svorts[1] = r - s2
        self.svorts[2] = self.diffx(q) - self.diffy(p)

    # ---- ENABLE_BUOYANCY: inversion_utils.f90:683-719, inversion.f90:232-292 ----
    def diffz(self, fs):
        """Spectral d/dz of a mixed-spectral field (inversion_utils.f90:683-719)."""
        nz = self.nz
        ds = fs[..., 0:1] * self.dphim + fs[..., nz:nz + 1] * self.dphip
        as_ = np.zeros_like(fs)
        as_[..., 1:nz] = self.rkz[None, None, 1:nz] * fs[..., 1:nz]
        ds = ds + self.fftcosine(as_)
        return self.field_decompose_semi_spectral(ds)

    def enable_buoyancy(self, buoy_phys, bfsq=0.0):
        """utils.f90:149-159 after the basic state has been removed: `buoy_phys` is b' in physical space."""
        self.buoyancy = True
        self.bfsq = float(bfsq)
        self.buoy = np.array(buoy_phys, dtype=np.float64)
        self.sbuoy = self.field_decompose_physical(self.buoy)
        self.sbuoys = np.zeros_like(self.sbuoy)
        self.bdiss = np.zeros((self.nx, self.ny))

    def init_diffusion_buoyancy(self, te, en, nnu=3, prediss=30.0, length_scale="Kolmogorov"):
        """inversion_utils.f90:144-151: bvisc, bhdis."""
        keep = (self.vvisc, self.nnu, self.vhdis)
        self.bvisc = self.init_diffusion(te, en, nnu, prediss, length_scale)
        self.bnnu, self.bhdis = self.nnu, self.vhdis
        self.vvisc, self.nnu, self.vhdis = keep
        return self.bvisc

    def buoyancy_tendency(self):
        """inversion.f90:232-292."""
        vel = self.vel
        self.buoy = self.field_combine_physical(self.sbuoy)
        fs = self.field_decompose_physical(vel[0] * self.buoy)
        btend = self.field_combine_physical(self.diffx(fs))
        fs = self.field_decompose_physical(vel[1] * self.buoy)
        fp = self.field_combine_physical(self.diffy(fs))
        btend = -btend - fp
        fs = self.field_decompose_physical(vel[2] * self.buoy)
        fp = self.field_combine_physical(self.diffz(fs))
        btend = btend - self.bfsq * vel[2] - fp
        self.sbuoys = self.field_decompose_physical(btend)

    def source(self):
        """inversion.f90:378-388."""
        if self.buoyancy:
            self.buoyancy_tendency()
        self.vorticity_tendency()

    def get_bfmax(self):
        """advance.f90:147-168."""
        sb = self.field_combine_semi_spectral(self.sbuoy)
        xp = self.fftxys2p(self.diffx(sb))
        yp = self.fftxys2p(self.diffy(sb))
        zp = self.fftxys2p(self.central_diffz(sb))
        xp = xp ** 2 + yp ** 2 + (zp + self.bfsq) ** 2
        return math.sqrt(math.sqrt(float(xp.max())))

    # ---- field_diagnostics.f90 ----
    def _trap(self, f):
        nz = self.nz
        return (np.sum(f[..., 1:nz]) + 0.5 * np.sum(f[..., 0]) + 0.5 * np.sum(f[..., nz]))

    def get_kinetic_energy(self):
        """field_diagnostics.f90:85-122."""
        return 0.5 * self._trap(self.vel[0] ** 2 + self.vel[1] ** 2 + self.vel[2] ** 2) * self.ncelli

    def get_enstrophy(self):
        """field_diagnostics.f90:172-206."""
        return 0.5 * self._trap(self.vor[0] ** 2 + self.vor[1] ** 2 + self.vor[2] ** 2) * self.ncelli

    def get_helicity(self):
        """plotting/nc_reader.py:94-101, plot_vor_vel_he_evolution.py:60-65."""
        h = self.vel[0] * self.vor[0] + self.vel[1] * self.vor[1] + self.vel[2] * self.vor[2]
        return self._trap(h) * self.ncelli

    def get_mean(self, ff):
        """field_diagnostics.f90:405-432."""
        return self._trap(ff) / float(self.ncell)

    def get_char_vorticity(self, vortrms):
        """field_diagnostics.f90:501-545."""
        vor = self.vor
        v1 = 0.5 * np.abs(vor[0][..., :-1] + vor[0][..., 1:])
        v2 = 0.5 * np.abs(vor[1][..., :-1] + vor[1][..., 1:])
        v3 = 0.5 * np.abs(vor[2][..., :-1] + vor[2][..., 1:])
        s = v1 + v2 + v3
        m = s > vortrms
        vorl1 = SMALL + np.sum(s[m])
        vorl2 = np.sum((v1 ** 2 + v2 ** 2 + v3 ** 2)[m])
        return vorl2 / vorl1

    def get_mean_vorticity(self):
        """field_diagnostics.f90:549-579."""
        return np.array([self._trap(self.vor[nc]) for nc in range(3)]) * self.ncelli

    def calc_vorticity_mean(self):
        """field_diagnostics.f90:584-599."""
        nz = self.nz
        savg = np.zeros(2)
        for nc in range(2):
            wk = np.zeros(nz)
            wk[: nz - 1] = self.svor[nc, 0, 0, 1:nz]
            wk = dst(wk, nz)
            savg[nc] = (0.5 * (self.svor[nc, 0, 0, 0] + self.svor[nc, 0, 0, nz])
                        + self.fnzi * np.sum(wk[: nz - 1]))
        return savg

    def adjust_vorticity_mean(self):
        """field_diagnostics.f90:604-619."""
        savg = self.calc_vorticity_mean()
        nz = self.nz
        for nc in range(2):
            self.svor[nc, 0, 0, 0] += self.ini_vor_mean[nc] - savg[nc]
            self.svor[nc, 0, 0, nz] += self.ini_vor_mean[nc] - savg[nc]

    def field_stats(self):
        """update_netcdf_field_diagnostics (field_diagnostics_netcdf.f90:257-439) plus the values adapt hands
        over with set_netcdf_field_diagnostic (advance.f90:188-193, 315-321, 366).  Needs vor2vel and adapt for
        the current state.  Keys: the reference's netCDF variable names, lower case."""
        nz = self.nz
        vor, vel = self.vor, self.vel
        d = self.diag
        ke = self.get_kinetic_energy()
        en = self.get_enstrophy()
        kexy = 0.5 * self._trap(vel[0] ** 2 + vel[1] ** 2) * self.ncelli          # field_diagnostics.f90:128-149
        enxy = 0.5 * self._trap(vor[0] ** 2 + vor[1] ** 2) * self.ncelli          # :211-229
        delta = self.horizontal_divergence()
        nxy = float(self.nx * self.ny)
        with np.errstate(divide="ignore", invalid="ignore"):
            romin = np.float64(vor[2].min()) / np.float64(self.f_cor[2])          # :293-304
            romax = np.float64(vor[2].max()) / np.float64(self.f_cor[2])          # :309-320
        return dict(
            ke=ke, en=en, omax=d["vortmax"], orms=d["vortrms"], ochar=d["vorch"],
            oxmean=d["vormean"][0], oymean=d["vormean"][1], ozmean=d["vormean"][2],
            kexy=kexy, kez=ke - kexy, enxy=enxy, enz=en - enxy,                   # :153-168, :248-263
            oxmin=vor[0].min(), oymin=vor[1].min(), ozmin=vor[2].min(),
            oxmax=vor[0].max(), oymax=vor[1].max(), ozmax=vor[2].max(),
            hemax=math.sqrt(np.max(vor[0] ** 2 + vor[1] ** 2)),                   # :233-244
            gmax=d["ggmax"], bfmax=d["bfmax"], umax=d["umax"], vmax=d["vmax"], wmax=d["wmax"],
            usoxmax=vor[0][..., nz].max(), lsoxmax=vor[0][..., 0].max(),
            usoymax=vor[1][..., nz].max(), lsoymax=vor[1][..., 0].max(),
            usozmax=vor[2][..., nz].max(), lsozmax=vor[2][..., 0].max(),
            usuhmax=math.sqrt(np.max(vel[0][..., nz] ** 2 + vel[1][..., nz] ** 2)),
            usgmax=d["usggmax"], lsgmax=d["lsggmax"],
            uszrms=math.sqrt(np.sum(vor[2][..., nz] ** 2) / nxy),
            usdelrms=math.sqrt(np.sum(delta[..., nz] ** 2) / nxy),
            rgmax=d["rmv"], rbfmax=0.0, rimin=0.0, romin=float(romin), romax=float(romax))

    def genspec(self):
        """genspec.f90:55-127: kinetic-energy spectrum of `vel` (needs vor2vel).  Returns (spec, num, dk);
        empty bins keep spec = 0 (the reference prints a warning, :118-119).  `kmag` as intended by :74-81 (the
        reference allocates it for a single x index, :34)."""
        nz = self.nz
        ke = self.get_kinetic_energy()
        s = [self.fftxyp2s(self.vel[nc]) for nc in range(3)]
        s[0] = dct(s[0], nz)
        s[1] = dct(s[1], nz)
        w = np.array(s[2])
        w[..., 1:] = dst(np.ascontiguousarray(w[..., 1:]), nz)
        s[2] = w
        kmag = np.floor(np.sqrt(self.rkx[:, None, None] ** 2 + self.rky[None, :, None] ** 2
                                + self.rkz[None, None, :] ** 2) + 0.5)                     # nint
        kmax = int(kmag.max())
        dk = kmax / math.sqrt((0.5 * self.nx) ** 2 + (0.5 * self.ny) ** 2 + float(nz) ** 2)
        m = (kmag * (1.0 / dk)).astype(np.int64)
        e = s[0] ** 2 + s[1] ** 2 + s[2] ** 2
        # the reference allocates spec(0:kmax) (:86-87) but int(kmag / dk) exceeds kmax when dk < 1 (boxes larger
        # than pi (2, 2, 1)): a latent out-of-bounds write there; bins here cover the largest reachable index
        nb = max(kmax, int(kmax * (1.0 / dk))) + 1
        spec = np.bincount(m.ravel(), weights=e.ravel(), minlength=nb).astype(np.float64)
        num = np.bincount(m.ravel(), minlength=nb).astype(np.float64)
        assert len(spec) == nb
        prefactor = 4.0 / 3.0 * math.pi * dk ** 3
        mm = np.arange(nb, dtype=np.float64)
        ok = num > 0
        spec[ok] = spec[ok] * prefactor * ((mm[ok] + 1) ** 3 - mm[ok] ** 3) / num[ok]
        spec *= ke / np.sum(spec * dk)
        return spec, num, dk

    # ---- fields_derived.f90:67-182 ----
    def pressure(self, dudx, dudy, dvdy, dwdx, dwdy):
        vor = self.vor
        dwdz = -(dudx + dvdy)
        pres = 2.0 * (dudx * dvdy - dudy * (vor[2] + dudy)
                      + dvdy * dwdz - dwdy * (dwdy - vor[0])
                      + dwdz * dudx - dwdx * (dwdx + vor[1]))
        if self.buoyancy:                                   # fields_derived.f90:108-112
            self.buoy = self.field_combine_physical(self.sbuoy)
            pres = pres + self.central_diffz(self.buoy) + self.f_cor[2] * vor[2]
        rs = self.fftxyp2s(pres)
        rs = dct(rs, self.nz)
        rs = self.green * rs
        rs = dct(rs, self.nz)
        return self.fftxys2p(rs)

    def horizontal_divergence(self):
        return self.fftxys2p(self.diffx(self.svel[0]) + self.diffy(self.svel[1]))

    # ---- utils.f90:136-184 setup_fields ----
    def set_vorticity(self, vor_phys, nnu=3, prediss=30.0, length_scale="Kolmogorov"):
        """Upload a physical vorticity field, decompose, record the initial
        mean, run vor2vel and initialise the (hyper)diffusion operator."""
        self.vor[:] = vor_phys
        for nc in range(3):
            self.svor[nc] = self.field_decompose_physical(self.vor[nc])
        self.ini_vor_mean = self.calc_vorticity_mean()
        self.vor2vel()
        ke = self.get_kinetic_energy()
        en = self.get_enstrophy()
        self.init_diffusion(ke, en, nnu, prediss, length_scale)
        return ke, en

    # ---- advance.f90:109-410 ----
    def strain_fields(self):
        """advance.f90:199-217: the five velocity-gradient fields."""
        dudx = self.fftxys2p(self.diffx(self.svel[0]))
        dudy = self.fftxys2p(self.diffy(self.svel[0]))
        dwdx = self.fftxys2p(self.diffx(self.svel[2]))
        dvdy = self.fftxys2p(self.diffy(self.svel[1]))
        dwdy = self.fftxys2p(self.diffy(self.svel[2]))
        return dudx, dudy, dvdy, dwdx, dwdy

    def adapt(self, t, time_limit, alpha=0.1, pretype="vorch", win=1000,
              with_pressure=False, bpretype=None, bwin=None):
        nz = self.nz
        vor, vel = self.vor, self.vel
        bfmax = self.get_bfmax() if self.buoyancy else 0.0
        xp = vor[0] ** 2 + vor[1] ** 2 + vor[2] ** 2
        vortmax = math.sqrt(np.max(np.abs(xp)))
        vortrms = math.sqrt(self.get_mean(xp))
        vorch = self.get_char_vorticity(vortrms)
        vormean = self.get_mean_vorticity()
        dudx, dudy, dvdy, dwdx, dwdy = self.strain_fields()
        s11 = dudx
        s12 = dudy + 0.5 * vor[2]
        s13 = dwdx + 0.5 * vor[1]
        s22 = dvdy
        s23 = dwdy - 0.5 * vor[0]
        s33 = -(dudx + dvdy)
        d1, d2, d3 = jacobi_eigenvalues(s11, s12, s13, s22, s23, s33)
        lmax = np.maximum(np.maximum(np.abs(d1), np.abs(d2)), np.abs(d3)).reshape(s11.shape)
        ggmax = max(np.finfo(np.float64).eps, float(lmax.max()))
        usggmax = max(0.0, float(lmax[..., nz].max()))
        lsggmax = max(0.0, float(lmax[..., 0].max()))
        if with_pressure:
            self.pres = self.pressure(dudx, dudy, dvdy, dwdx, dwdy)
            self.delta = self.horizontal_divergence()
        umax = float(vel[0].max())
        vmax = float(vel[1].max())
        wmax = float(vel[2].max())
        dtcfl = CFLMAX * min(self.dx[0] / (umax + SMALL),
                             self.dx[1] / (vmax + SMALL),
                             self.dx[2] / (wmax + SMALL))
        dt = min(alpha / (ggmax + SMALL), alpha / (bfmax + SMALL), dtcfl, time_limit - t)
        rmb = 0.0
        if self.buoyancy:                                   # advance.f90:358-362
            if self.buoy_rollmean is None:
                self.buoy_rollmean = RollingMean(bwin if bwin else win)
            rmb = self.buoy_rollmean.get_next(bfmax)
        if self.rollmean is None:
            self.rollmean = RollingMean(win)
        rmv = self.rollmean.get_next(ggmax)
        table = {"constant": 1.0, "vorch": vorch, "bfmax": bfmax,
                 "roll-mean-max-strain": rmv, "roll-mean-bfmax": rmb, "max-strain": ggmax,
                 "us-max-strain": usggmax}
        pref = table[pretype]
        self.bpref = table[bpretype if bpretype else pretype]  # bval = get_diffusion_pre_factor(buoy_visc) (:372-374)
        self.rmb = rmb
        self.diag = dict(vortmax=vortmax, vortrms=vortrms, vorch=vorch,
                         vormean=vormean, bfmax=bfmax, ggmax=ggmax, umax=umax,
                         vmax=vmax, wmax=wmax, usggmax=usggmax, lsggmax=lsggmax,
                         rmv=rmv, dt=dt, pref=pref)
        return dt, pref

    # ---- cn2.f90 ----
    def cn2_set_diffusion(self, dt, vorch, bf=0.0):
        dfac = dt if self.nnu == 1 else vorch * dt
        self.vdiss = 1.0 / (1.0 + dfac * self.vhdis)
        if self.buoyancy:                                    # cn2.f90:61-75
            dbac = dt if self.bnnu == 1 else bf * dt
            self.bdiss = 1.0 / (1.0 + dbac * self.bhdis)

    def _cn2_update_buoy(self, dt2, literal):
        """cn2.f90:107-117 / :151-160."""
        q = self.filt * (self.bsm + dt2 * self.sbuoys)
        if literal:
            q = self.field_combine_semi_spectral(q)
            q = self.bdiss[..., None] * q
            q = self.field_decompose_semi_spectral(q)
        else:
            q = self.bdiss[..., None] * q
        self.sbuoy = q

    def _cn2_update(self, dt2, literal):
        vd = self.vdiss[..., None]
        for nc in range(3):
            q = self.filt * (self.vortsm[nc] + dt2 * self.svorts[nc])
            if literal:
                q = self.field_combine_semi_spectral(q)
                q = vd * q
                q = self.field_decompose_semi_spectral(q)
            else:
                q = vd * q
            self.svor[nc] = q
        self.adjust_vorticity_mean()

    def cn2_step(self, t, dt, literal=True, niter=2):
        """cn2.f90:92-181.  literal=False uses the identity
        decompose(c(ky,kx)*combine(q)) == c*q."""
        dt2 = 0.5 * dt
        if self.buoyancy:
            self.bsm = self.sbuoy + dt2 * self.sbuoys
            self._cn2_update_buoy(dt2, literal)
        self.vortsm = self.svor + dt2 * self.svorts
        self._cn2_update(dt2, literal)
        for _ in range(niter):
            self.vor2vel()
            self.source()
            if self.buoyancy:
                self._cn2_update_buoy(dt2, literal)
            self._cn2_update(dt2, literal)
        return t + dt

    # ---- impl_rk4.f90 ----
    def rk4_set_diffusion(self, dt, vorch, bf=0.0):
        self.vdiss = 0.5 * vorch * dt * self.vhdis
        if self.buoyancy:                                    # impl_rk4.f90:44-50
            self.bdiss = 0.5 * bf * dt * self.bhdis

    def _cmd(self, q, fac, literal):
        """combine -> multiply by fac(ky,kx) -> decompose."""
        if literal:
            q = self.field_combine_semi_spectral(q)
            q = fac[..., None] * q
            return self.field_decompose_semi_spectral(q)
        return fac[..., None] * q

    def rk4_step(self, t, dt, literal=True):
        """impl_rk4.f90:76-207."""
        dt2, dt3, dt6 = 0.5 * dt, dt / 3.0, dt / 6.0
        epq = np.exp(self.vdiss)
        emq = 1.0 / epq
        epq = epq * self.filt[..., 0]
        svor, svorts = self.svor, self.svorts
        svori = np.empty_like(svor)
        svorf = np.empty_like(svor)
        f0 = self.filt[..., 0:1]
        B = self.buoyancy
        if B:                                                 # impl_rk4.f90:91-105
            bpq = np.exp(self.bdiss)
            bmq = 1.0 / bpq
            bpq = bpq * self.filt[..., 0]
            self.sbuoys = f0 * self.sbuoys
            sbuoyi = self.sbuoy
            self.sbuoy = self._cmd(sbuoyi + dt2 * self.sbuoys, bmq, literal)
            sbuoyf = sbuoyi + dt6 * self.sbuoys
        for nc in range(3):                                   # substep one
            svorts[nc] = f0 * svorts[nc]
            svori[nc] = svor[nc]
            svor[nc] = self._cmd(svori[nc] + dt2 * svorts[nc], emq, literal)
            svorf[nc] = svori[nc] + dt6 * svorts[nc]
        self.vor2vel(); self.source()
        t = t + dt2
        if B:                                                 # :124-131
            self.sbuoys = self._cmd(self.sbuoys, bpq, literal)
            self.sbuoy = self._cmd(sbuoyi + dt2 * self.sbuoys, bmq, literal)
            sbuoyf = sbuoyf + dt3 * self.sbuoys
        for nc in range(3):                                   # substep two
            svorts[nc] = self._cmd(svorts[nc], epq, literal)
            svor[nc] = self._cmd(svori[nc] + dt2 * svorts[nc], emq, literal)
            svorf[nc] = svorf[nc] + dt3 * svorts[nc]
        self.vor2vel(); self.source()
        t = t + dt2
        emq = emq ** 2
        if B:                                                 # :154-164
            bmq = bmq ** 2
            self.sbuoys = self._cmd(self.sbuoys, bpq, literal)
            self.sbuoy = self._cmd(sbuoyi + dt * self.sbuoys, bmq, literal)
            sbuoyf = sbuoyf + dt3 * self.sbuoys
        for nc in range(3):                                   # substep three
            svorts[nc] = self._cmd(svorts[nc], epq, literal)
            svor[nc] = self._cmd(svori[nc] + dt * svorts[nc], emq, literal)
            svorf[nc] = svorf[nc] + dt3 * svorts[nc]
        self.vor2vel(); self.source()
        epq = epq ** 2
        if B:                                                 # :187-195
            bpq = bpq ** 2
            self.sbuoys = self._cmd(self.sbuoys, bpq, literal)
            self.sbuoy = self._cmd(sbuoyf + dt6 * self.sbuoys, bmq, literal)
        for nc in range(3):                                   # substep four
            svorts[nc] = self._cmd(svorts[nc], epq, literal)
            svor[nc] = self._cmd(svorf[nc] + dt6 * svorts[nc], emq, literal)
        self.adjust_vorticity_mean()
        return t

    # ---- advance.f90:77-104 ----
    def advance(self, t, time_limit, stepper="cn2", alpha=0.1, pretype="vorch",
                win=1000, literal=True, with_pressure=False, bpretype=None, bwin=None):
        self.vor2vel()
        dt, pref = self.adapt(t, time_limit, alpha, pretype, win, with_pressure, bpretype, bwin)
        if stepper == "cn2":
            self.cn2_set_diffusion(dt, pref, self.bpref)
        else:
            self.rk4_set_diffusion(dt, pref, self.bpref)
        self.source()
        if stepper == "cn2":
            return self.cn2_step(t, dt, literal), dt
        return self.rk4_step(t, dt, literal), dt


# --------------------------------------------------------------------------
# Initial conditions
# --------------------------------------------------------------------------
def beltrami_vorticity(nx, ny, nz, lower, extent, k=2, l=2, m=1):
    """beltrami.f90:141-181 (`beltrami_init`, `get_flow_vorticity`)."""
    lower = np.asarray(lower, dtype=np.float64)
    extent = np.asarray(extent, dtype=np.float64)
    dx = extent / np.array([nx, ny, nz], dtype=np.float64)
    kk, ll, mm = float(k), float(l), float(m)
    alpha = math.sqrt(kk ** 2 + ll ** 2 + mm ** 2)
    fk2l2 = alpha / float(k ** 2 + l ** 2)
    x = (lower[0] + dx[0] * np.arange(nx))[:, None, None]
    y = (lower[1] + dx[1] * np.arange(ny))[None, :, None]
    z = (lower[2] + dx[2] * np.arange(nz + 1))[None, None, :]
    cosmz, sinmz = np.cos(mm * z), np.sin(mm * z)
    s, c = np.sin(kk * x + ll * y), np.cos(kk * x + ll * y)
    vor = np.empty((3, nx, ny, nz + 1))
    vor[0] = fk2l2 * (kk * mm * sinmz - ll * alpha * cosmz) * s
    vor[1] = fk2l2 * (ll * mm * sinmz + kk * alpha * cosmz) * s
    vor[2] = alpha * cosmz * c
    return vor


def beltrami_setup(n, stepper="cn2", **kw):
    """examples/beltrami_<n>.config + beltrami<n>x<n>x<n>.nml."""
    lower = -0.5 * math.pi * np.ones(3)
    extent = math.pi * np.ones(3)
    s = PS3D(n, n, n, lower, extent, filtering="Hou & Li")
    vor = beltrami_vorticity(n, n, n, lower, extent)
    s.set_vorticity(vor, nnu=3, prediss=30.0, length_scale="Kolmogorov")
    return s
