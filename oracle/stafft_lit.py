"""ctypes wrapper of oracle/stafft_lit.c, the literal C restatement of the reference's src/fft/stafft.f90.

TEST INFRASTRUCTURE ONLY: imported by tests/ (pins oracle/ps3d_oracle.py's library transforms and the CUDA
transforms to the arithmetic of the Fortran build, measures the reference's own round-off floor).
Build: `python -c "import __graft_entry__ as g; g.build_stafft_lit()"` (gcc -O2 -ffp-contract=off: no FMA
contraction, the operations stay the ones written in the Fortran source)."""
import ctypes as C
import os

import numpy as np

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libstafft_lit.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _p(a):
    return a.ctypes.data_as(_dp)


class Stafft:
    """initfft(n) once, then forfft / revfft / dct / dst on arrays whose LAST axis is the transform axis
    (each line is one call with m = 1, as sta3dfft.f90:161-178 and inversion_utils.f90:573-590 do)."""

    def __init__(self, n, path=LIB_PATH):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path}: build it with __graft_entry__.build_stafft_lit()")
        self.dll = C.CDLL(path)
        for name in ("lit_forfft", "lit_revfft", "lit_dct", "lit_dst"):
            getattr(self.dll, name).argtypes = [C.c_int, C.c_int, _dp, _dp, _ip]
        self.dll.lit_initfft.argtypes = [C.c_int, _ip, _dp]
        self.n = n
        self.factors = np.zeros(5, dtype=np.int32)
        self.trig = np.zeros(2 * n)
        rc = self.dll.lit_initfft(n, self.factors.ctypes.data_as(_ip), _p(self.trig))
        if rc:
            raise ValueError(f"stafft_lit: length {n} " + ("cannot be factorised (stafft.f90:87-95)" if rc == 1 else
                                                          "needs radix 5 or 6, not restated"))

    def _lines(self, name, x, width):
        x = np.array(x, dtype=np.float64, order="C", copy=True)
        assert x.shape[-1] == width
        flat = x.reshape(-1, width)
        fn = getattr(self.dll, name)
        fp = self.factors.ctypes.data_as(_ip)
        for row in flat:                 # m = 1 per call
            rc = fn(1, self.n, _p(row), _p(self.trig), fp)
            assert rc == 0
        return x

    def forfft(self, x): return self._lines("lit_forfft", x, self.n)
    def revfft(self, x): return self._lines("lit_revfft", x, self.n)
    def dct(self, x): return self._lines("lit_dct", x, self.n + 1)       # x(0:n)
    def dst(self, x): return self._lines("lit_dst", x, self.n)           # x(1:n)

    def forfft_m(self, x):
        """m > 1 in one call: x has shape (n, m) in C order == Fortran x(m, n), vector index fastest."""
        x = np.array(x, dtype=np.float64, order="C", copy=True)
        n, m = x.shape
        assert n == self.n
        assert self.dll.lit_forfft(m, n, _p(x), _p(self.trig), self.factors.ctypes.data_as(_ip)) == 0
        return x
