"""ctypes wrapper of oracle/ps3d_ref.cpp, the C++/OpenMP restatement of the reference's time-step path.

TEST / BASELINE INFRASTRUCTURE: imported only by tests/ (cross-check against ps3d_oracle.py) and by bench.py's
CPU legs.  Build with `python -c "import __graft_entry__ as g; g.build_ref()"`."""
import ctypes as C
import os

import numpy as np

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libps3d_ref.so")
_dp = C.POINTER(C.c_double)
FIELDS = {"svor": 0, "vor": 1, "vel": 2, "svel": 3, "svorts": 4}
OPS = {"fftxyp2s": 0, "fftxys2p": 1, "fftsine": 2, "fftcosine": 3, "diffx": 4, "diffy": 5, "central_diffz": 6,
       "field_combine_semi_spectral": 7, "field_decompose_semi_spectral": 8}


def _p(a):
    return a.ctypes.data_as(_dp)


class RefSolver:
    """One-rank solver state of the restatement (power-of-two grids, cn2 / impl-diff-rk4, pretype 'vorch', Kolmogorov scaling)."""

    def __init__(self, nx, ny, nz, lower, extent, filtering="Hou & Li", path=LIB_PATH):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path}: build it with __graft_entry__.build_ref()")
        self.dll = C.CDLL(path)
        d = self.dll
        d.ps3d_ref_create.restype = C.c_void_p
        d.ps3d_ref_create.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_int]
        d.ps3d_ref_destroy.argtypes = [C.c_void_p]
        d.ps3d_ref_set_vorticity.argtypes = [C.c_void_p, _dp, C.c_int, C.c_double, _dp]
        d.ps3d_ref_advance.restype = C.c_double
        d.ps3d_ref_advance.argtypes = [C.c_void_p, _dp, C.c_double, C.c_double, C.c_int]
        d.ps3d_ref_get.argtypes = [C.c_void_p, C.c_int, _dp]
        d.ps3d_ref_vor2vel.argtypes = [C.c_void_p]
        d.ps3d_ref_op.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        d.ps3d_ref_diag.argtypes = [C.c_void_p, _dp]
        d.ps3d_ref_set_threads.argtypes = [C.c_int]
        d.ps3d_ref_set_threads.restype = C.c_int
        # all host cores, whatever OMP_NUM_THREADS a launcher exported (torchrun sets it to 1)
        self.threads = d.ps3d_ref_set_threads(os.cpu_count() or 0)
        self.shape = (nx, ny, nz + 1)
        lo = np.ascontiguousarray(lower, dtype=np.float64)
        ex = np.ascontiguousarray(extent, dtype=np.float64)
        self.h = d.ps3d_ref_create(nx, ny, nz, _p(lo), _p(ex), 0 if filtering == "Hou & Li" else 1)
        if not self.h:
            raise ValueError("ps3d_ref: power-of-two grids >= 8 only")
        self.t = np.zeros(1)

    def set_vorticity(self, vor, nnu=3, prediss=30.0):
        v = np.ascontiguousarray(vor, dtype=np.float64)
        out = np.zeros(2)
        self.dll.ps3d_ref_set_vorticity(self.h, _p(v), nnu, prediss, _p(out))
        return float(out[0]), float(out[1])

    def advance(self, time_limit=100.0, alpha=0.1, stepper="cn2"):
        dt = self.dll.ps3d_ref_advance(self.h, _p(self.t), time_limit, alpha, 0 if stepper == "cn2" else 1)
        return float(self.t[0]), float(dt)

    def vor2vel(self):
        self.dll.ps3d_ref_vor2vel(self.h)

    def get(self, name):
        out = np.empty((3,) + self.shape)
        self.dll.ps3d_ref_get(self.h, FIELDS[name], _p(out))
        return out

    def op(self, name, f):
        a = np.ascontiguousarray(f, dtype=np.float64)
        out = np.empty_like(a)
        self.dll.ps3d_ref_op(self.h, OPS[name], _p(a), _p(out))
        return out

    def diag(self):
        out = np.zeros(2)
        self.dll.ps3d_ref_diag(self.h, _p(out))
        return dict(vorch=float(out[0]), ggmax=float(out[1]))

    def close(self):
        if self.h:
            self.dll.ps3d_ref_destroy(self.h)
            self.h = None
