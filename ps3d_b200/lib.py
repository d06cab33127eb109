"""ctypes binding of the C ABI in include/ps3d_cuda.h (one function per entry point)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libps3d_cuda.so")

FILTER = {"Hou & Li": 0, "2/3-rule": 1}
LSCALE = {"Kolmogorov": 0, "geophysical": 1}
STEPPER = {"cn2": 0, "impl-diff-rk4": 1}
PRETYPE = {"constant": 0, "vorch": 1, "bfmax": 2, "roll-mean-max-strain": 3, "max-strain": 4, "us-max-strain": 5,
           "roll-mean-bfmax": 6}
FIELD = {"svor": 0, "vor": 1, "vel": 2, "svel": 3, "svorts": 4, "pres": 5, "delta": 6, "sbuoy": 7, "buoy": 8,
         "sbuoys": 9}
DIAG = ["vortmax", "vortrms", "vorch", "vormean_x", "vormean_y", "vormean_z", "bfmax", "ggmax", "umax", "vmax",
        "wmax", "usggmax", "lsggmax", "rmv", "dt", "prefactor"]

_dp = C.POINTER(C.c_double)

# name -> argtypes; every function returns int except the introspection ones
# field_diagnostics_netcdf.f90:36-75 (NC_KE = 1 ... NC_ROMAX = 40), in order
NC_NAMES = ("ke", "en", "omax", "orms", "ochar", "oxmean", "oymean", "ozmean", "kexy", "kez", "enxy", "enz",
            "oxmin", "oymin", "ozmin", "oxmax", "oymax", "ozmax", "hemax", "gmax", "bfmax", "umax", "vmax", "wmax",
            "usoxmax", "lsoxmax", "usoymax", "lsoymax", "usozmax", "lsozmax", "usuhmax", "usgmax", "lsgmax",
            "uszrms", "usdelrms", "rgmax", "rbfmax", "rimin", "romin", "romax")

_SIGNATURES = {
    "ps3d_cuda_init": [C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_int, C.c_void_p],
    "ps3d_cuda_init_inversion": [C.c_int],
    "ps3d_cuda_init_diffusion": [C.c_int, C.c_double, C.c_int, C.c_double, C.c_double, _dp],
    "ps3d_cuda_finalise": [],
    "ps3d_cuda_field_stats": [_dp],
    "ps3d_cuda_genspec": [C.c_int, _dp, _dp, C.POINTER(C.c_int), _dp],
    "ps3d_cuda_fftxyp2s": [_dp, _dp],
    "ps3d_cuda_fftxys2p": [_dp, _dp],
    "ps3d_cuda_fftsine": [_dp],
    "ps3d_cuda_fftcosine": [_dp],
    "ps3d_cuda_diffx": [_dp, _dp],
    "ps3d_cuda_diffy": [_dp, _dp],
    "ps3d_cuda_central_diffz": [_dp, _dp],
    "ps3d_cuda_diffz": [_dp, _dp],
    "ps3d_cuda_field_combine_semi_spectral": [_dp],
    "ps3d_cuda_field_decompose_semi_spectral": [_dp],
    "ps3d_cuda_field_combine_physical": [_dp, _dp],
    "ps3d_cuda_field_decompose_physical": [_dp, _dp],
    "ps3d_cuda_upload_vorticity": [_dp],
    "ps3d_cuda_upload_vorticity_begin": [_dp],
    "ps3d_cuda_upload_vorticity_end": [],
    "ps3d_cuda_vor2vel": [],
    "ps3d_cuda_source": [],
    "ps3d_cuda_adapt": [C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, _dp, _dp],
    "ps3d_cuda_stepper_setup": [C.c_int],
    "ps3d_cuda_set_diffusion": [C.c_double, C.c_double],
    "ps3d_cuda_step": [_dp, C.c_double],
    "ps3d_cuda_advance": [_dp, C.c_double, C.c_double, C.c_int, C.c_int, _dp, _dp],
    "ps3d_cuda_download": [C.c_int, C.c_int, _dp],
    "ps3d_cuda_upload": [C.c_int, C.c_int, _dp],
    "ps3d_cuda_diagnostics": [_dp],
    "ps3d_cuda_time_kernel": [C.c_int, C.c_int, _dp],
    "ps3d_cuda_set_transport": [C.c_void_p, C.c_void_p, C.c_void_p],
    "ps3d_cuda_comm_stats": [C.POINTER(C.c_longlong), _dp],
    "ps3d_cuda_set_physics": [_dp, C.c_double],
    "ps3d_cuda_enable_buoyancy": [],
    "ps3d_cuda_upload_buoyancy": [_dp],
    "ps3d_cuda_init_diffusion_buoyancy": [C.c_int, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, _dp],
    "ps3d_cuda_set_diffusion_buoyancy": [C.c_double, C.c_double],
    "ps3d_cuda_buoyancy_diag": [_dp],
}
ALLTOALL_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, _dp, C.c_int, C.c_int, C.c_void_p)
SPECTRAL_FIELDS = ("svor", "svel", "svorts", "sbuoy", "sbuoys")
EXPORTED_SYMBOLS = sorted(list(_SIGNATURES) + ["ps3d_cuda_last_error", "ps3d_cuda_kernel_launches",
                                               "ps3d_cuda_tma_launches", "ps3d_cuda_last_advance_ms"])


class PS3DError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"libps3d_cuda status {status}: {message}")
        self.status = status


def _ptr(a):
    return a.ctypes.data_as(_dp)


def _in(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class PS3DLib:
    """Thin, explicit binding.  Arrays are (nx, ny, nz+1) C-order float64, i.e. the
    memory image of the reference's Fortran `f(0:nz, 0:ny-1, 0:nx-1)`."""

    def __init__(self, path=LIB_PATH):
        if not os.path.exists(path):
            raise PS3DError(-1, f"{path} not found: build it with `python -c 'import __graft_entry__ as g; "
                                f"g.build()'` (there is no CPU fallback)")
        self.path = path
        self.dll = C.CDLL(path)
        for name, args in _SIGNATURES.items():
            fn = getattr(self.dll, name)
            fn.argtypes = args
            fn.restype = C.c_int
        self.dll.ps3d_cuda_last_error.restype = C.c_char_p
        self.dll.ps3d_cuda_kernel_launches.restype = C.c_longlong
        self.dll.ps3d_cuda_tma_launches.restype = C.c_longlong
        self.dll.ps3d_cuda_last_advance_ms.restype = C.c_double
        self.shape = self.spec_shape = None

    def _call(self, name, *args):
        st = getattr(self.dll, name)(*args)
        if st != 0:
            raise PS3DError(st, self.dll.ps3d_cuda_last_error().decode())

    # ---- setup ----
    def init(self, nx, ny, nz, lower, extent, rank=0, nranks=1, nccl_id=None):
        lo = _in(lower)
        ex = _in(extent)
        idbuf = None
        if nccl_id is not None:
            assert len(nccl_id) == 128, "ncclUniqueId is 128 bytes"
            idbuf = C.cast(C.create_string_buffer(bytes(nccl_id), 128), C.c_void_p)
        self._call("ps3d_cuda_init", nx, ny, nz, _ptr(lo), _ptr(ex), rank, nranks, idbuf)
        self.shape = (nx // nranks, ny, nz + 1)                      # physical fields: x-slab
        # spectral fields: all kx, natural ky on one rank / this rank's slab of the paired ky order otherwise
        self.spec_shape = (nx, ny // nranks, nz + 1) if nranks > 1 else self.shape
        self.nranks, self.rank = nranks, rank

    def init_inversion(self, filtering="Hou & Li"):
        self._call("ps3d_cuda_init_inversion", FILTER[filtering])

    def init_diffusion(self, te, en, nnu=3, prediss=30.0, length_scale="Kolmogorov"):
        nu = C.c_double(0.0)
        self._call("ps3d_cuda_init_diffusion", nnu, prediss, LSCALE[length_scale], te, en, C.byref(nu))
        return nu.value

    # ---- physics.f90 / the ENABLE_BUOYANCY build ----
    def set_physics(self, f_cor=(0.0, 0.0, 0.0), bfsq=0.0):
        fc = _in(f_cor)
        self._call("ps3d_cuda_set_physics", _ptr(fc), bfsq)

    def enable_buoyancy(self): self._call("ps3d_cuda_enable_buoyancy")

    def upload_buoyancy(self, buoy):
        buoy = _in(buoy)
        assert buoy.shape == self.shape, (buoy.shape, self.shape)
        self._call("ps3d_cuda_upload_buoyancy", _ptr(buoy))

    def init_diffusion_buoyancy(self, te, en, nnu=3, prediss=30.0, length_scale="Kolmogorov", pretype="vorch", win=1000):
        nu = C.c_double(0.0)
        self._call("ps3d_cuda_init_diffusion_buoyancy", nnu, prediss, LSCALE[length_scale], te, en, PRETYPE[pretype], win,
                   C.byref(nu))
        return nu.value

    def set_diffusion_buoyancy(self, dt, bf): self._call("ps3d_cuda_set_diffusion_buoyancy", dt, bf)

    def buoyancy_diag(self):
        out = np.zeros(4)
        self._call("ps3d_cuda_buoyancy_diag", _ptr(out))
        return dict(bfmax=out[0], rmb=out[1], bval=out[2], bvisc=out[3])

    def finalise(self):
        self._call("ps3d_cuda_finalise")
        self.shape = None

    # ---- operator mode ----
    def _op2(self, name, a):
        a = _in(a)
        out = np.empty_like(a)
        self._call(name, _ptr(a), _ptr(out))
        return out

    def _op1(self, name, a):
        a = np.array(a, dtype=np.float64, order="C", copy=True)
        self._call(name, _ptr(a))
        return a

    def fftxyp2s(self, fp): return self._op2("ps3d_cuda_fftxyp2s", fp)
    def fftxys2p(self, fs): return self._op2("ps3d_cuda_fftxys2p", fs)
    def fftsine(self, fs): return self._op1("ps3d_cuda_fftsine", fs)
    def fftcosine(self, fs): return self._op1("ps3d_cuda_fftcosine", fs)
    def diffx(self, fs): return self._op2("ps3d_cuda_diffx", fs)
    def diffy(self, fs): return self._op2("ps3d_cuda_diffy", fs)
    def central_diffz(self, fs): return self._op2("ps3d_cuda_central_diffz", fs)
    def diffz(self, fs): return self._op2("ps3d_cuda_diffz", fs)
    def field_combine_semi_spectral(self, sf): return self._op1("ps3d_cuda_field_combine_semi_spectral", sf)
    def field_decompose_semi_spectral(self, sf): return self._op1("ps3d_cuda_field_decompose_semi_spectral", sf)
    def field_combine_physical(self, sf): return self._op2("ps3d_cuda_field_combine_physical", sf)
    def field_decompose_physical(self, fc): return self._op2("ps3d_cuda_field_decompose_physical", fc)

    # ---- resident mode ----
    def upload_vorticity(self, vor):
        vor = _in(vor)
        assert vor.shape == (3,) + self.shape, (vor.shape, self.shape)
        self._call("ps3d_cuda_upload_vorticity", _ptr(vor))

    def upload_vorticity_begin(self, vor):
        """Queue the host -> device copies and return; `vor` (C-contiguous float64, ideally pinned) must stay alive
        until upload_vorticity_end()."""
        assert vor.dtype == np.float64 and vor.flags["C_CONTIGUOUS"] and vor.shape == (3,) + self.shape
        self._call("ps3d_cuda_upload_vorticity_begin", _ptr(vor))
        self._pending_uploads = getattr(self, "_pending_uploads", []) + [vor]      # keep the buffers alive (FIFO, depth <= 2)

    def upload_vorticity_end(self):
        self._call("ps3d_cuda_upload_vorticity_end")
        self._pending_uploads = getattr(self, "_pending_uploads", [None])[1:]

    def vor2vel(self): self._call("ps3d_cuda_vor2vel")
    def source(self): self._call("ps3d_cuda_source")

    def adapt(self, t, t_limit, alpha=0.1, pretype="vorch", win=1000):
        dt = C.c_double(0.0)
        diag = np.zeros(16)
        self._call("ps3d_cuda_adapt", t, t_limit, alpha, PRETYPE[pretype], win, C.byref(dt), _ptr(diag))
        return dt.value, dict(zip(DIAG, diag))

    def stepper_setup(self, stepper="cn2"): self._call("ps3d_cuda_stepper_setup", STEPPER[stepper])
    def set_diffusion(self, dt, pref): self._call("ps3d_cuda_set_diffusion", dt, pref)

    def step(self, t, dt):
        tt = C.c_double(t)
        self._call("ps3d_cuda_step", C.byref(tt), dt)
        return tt.value

    def advance(self, t, t_limit, alpha=0.1, pretype="vorch", win=1000):
        tt = C.c_double(t)
        dt = C.c_double(0.0)
        diag = np.zeros(16)
        self._call("ps3d_cuda_advance", C.byref(tt), t_limit, alpha, PRETYPE[pretype], win, C.byref(dt), _ptr(diag))
        return tt.value, dt.value, dict(zip(DIAG, diag))

    def set_transport(self, alltoall, allreduce):
        """Plug host collectives (ctypes callbacks built with ALLTOALL_FN / ALLREDUCE_FN); keep them alive."""
        self._cb = (alltoall, allreduce)
        self._call("ps3d_cuda_set_transport", C.cast(alltoall, C.c_void_p), C.cast(allreduce, C.c_void_p), None)

    def comm_stats(self):
        n = C.c_longlong(0)
        b = C.c_double(0.0)
        self._call("ps3d_cuda_comm_stats", C.byref(n), C.byref(b))
        return n.value, b.value

    @staticmethod
    def paired_ky(ny):
        """ky of each row of the paired order ky' = 0, ny/2, 1, ny-1, 2, ny-2, ..."""
        out = np.empty(ny, dtype=np.int64)
        out[0], out[1] = 0, ny // 2
        for a in range(1, ny // 2):
            out[2 * a], out[2 * a + 1] = a, ny - a
        return out

    def download(self, field, comp=0):
        out = np.empty(self.spec_shape if field in SPECTRAL_FIELDS else self.shape)
        self._call("ps3d_cuda_download", FIELD[field], comp, _ptr(out))
        return out

    def pressure(self): return self.download("pres")
    def horizontal_divergence(self): return self.download("delta")

    def download3(self, field):
        return np.stack([self.download(field, c) for c in range(3)])

    def upload(self, field, comp, a):
        a = _in(a)
        self._call("ps3d_cuda_upload", FIELD[field], comp, _ptr(a))

    def diagnostics(self):
        out = np.zeros(8)
        self._call("ps3d_cuda_diagnostics", _ptr(out))
        return dict(ke=out[0], en=out[1], helicity=out[2], hke=out[3], vke=out[4], hen=out[5], ven=out[6], hemax=out[7])

    def field_stats(self):
        """The 40 scalars of the field-statistics file (field_diagnostics_netcdf.f90:36-75), keyed by the
        reference's netCDF names in lower case; needs vor2vel + adapt for the current state."""
        out = np.zeros(len(NC_NAMES))
        self._call("ps3d_cuda_field_stats", _ptr(out))
        return dict(zip(NC_NAMES, out.tolist()))

    def genspec(self):
        """Kinetic-energy spectrum of the current velocity (genspec.f90): returns (spec, num, dk)."""
        nb = C.c_int(0)
        dk = np.zeros(1)
        self._call("ps3d_cuda_genspec", 0, None, None, C.byref(nb), _ptr(dk))
        spec, num = np.zeros(nb.value), np.zeros(nb.value)
        self._call("ps3d_cuda_genspec", nb.value, _ptr(spec), _ptr(num), C.byref(nb), _ptr(dk))
        return spec, num, float(dk[0])

    def kernel_launches(self): return int(self.dll.ps3d_cuda_kernel_launches())
    def tma_launches(self): return int(self.dll.ps3d_cuda_tma_launches())
    def last_advance_ms(self): return float(self.dll.ps3d_cuda_last_advance_ms())

    def time_kernel(self, which, reps):
        ms = C.c_double(0.0)
        self._call("ps3d_cuda_time_kernel", which, reps, C.byref(ms))
        return ms.value


_lib = None


def load():
    """The product library (CUDA).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        _lib = PS3DLib(LIB_PATH)
    return _lib
