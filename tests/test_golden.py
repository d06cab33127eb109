"""Committed golden vectors (tests/golden, made by tests/golden/make_golden.py from the oracle): the oracle must
still reproduce them (CPU), and the CUDA path must match them (GPU)."""
import math
import os

import numpy as np
import pytest

from oracle import ps3d_oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
COLS = ["t", "dt", "vortmax", "vortrms", "vorch", "ggmax", "umax", "vmax", "wmax", "usggmax", "lsggmax"]


def test_oracle_reproduces_golden_operators():
    g = np.load(os.path.join(G, "operators_16x32x8.npz"))
    s = O.PS3D(16, 32, 8, g["lower"], g["extent"])
    f = g["f"]
    for name, fn in (("fftxyp2s", s.fftxyp2s), ("fftsine", s.fftsine), ("fftcosine", s.fftcosine), ("diffx", s.diffx),
                     ("diffy", s.diffy), ("diffz", s.central_diffz), ("combine", s.field_combine_semi_spectral),
                     ("decompose", s.field_decompose_semi_spectral)):
        assert np.max(np.abs(fn(f) - g[name])) < 1e-13, name


def test_oracle_reproduces_golden_trajectory_prefix():
    g = np.load(os.path.join(G, "beltrami32_cn2_100steps.npz"))
    s = O.beltrami_setup(32)
    t = 0.0
    for i in range(5):
        t, dt = s.advance(t, 100.0, "cn2", literal=True)
        assert t == pytest.approx(g["series"][i, 0], rel=1e-12) and dt == pytest.approx(g["series"][i, 1], rel=1e-12)
        assert s.diag["vorch"] == pytest.approx(g["series"][i, 4], rel=1e-11)


@pytest.mark.parametrize("stepper,tag", [("cn2", "cn2"), ("impl-diff-rk4", "rk4")])
def test_cpp_restatement_reproduces_golden_trajectory(stepper, tag):
    """The independent C++/OpenMP restatement (oracle/ps3d_ref.cpp: its own FFTs, the reference's transposes and
    literal steppers) reproduces the committed 100-step series: t, dt, vorch and ggmax of every step."""
    import __graft_entry__ as GE
    from oracle.ps3d_ref import RefSolver
    GE.build_ref()
    g = np.load(os.path.join(G, f"beltrami32_{tag}_100steps.npz"))
    lower = -0.5 * math.pi * np.ones(3)
    extent = math.pi * np.ones(3)
    r = RefSolver(32, 32, 32, lower, extent)
    try:
        r.set_vorticity(O.beltrami_vorticity(32, 32, 32, lower, extent))
        for i in range(100):
            t, dt = r.advance(stepper=stepper)
            row = dict(zip(COLS, g["series"][i]))
            assert t == pytest.approx(row["t"], rel=1e-11) and dt == pytest.approx(row["dt"], rel=1e-11), i
            d = r.diag()
            assert d["vorch"] == pytest.approx(row["vorch"], rel=1e-10) and d["ggmax"] == pytest.approx(row["ggmax"], rel=1e-10), i
        r.vor2vel()                      # the sample was taken after vor2vel (make_golden.py)
        svor = r.get("svor")[:, ::4, ::4, ::4]
        assert np.max(np.abs(svor - g["svor_sample"])) < 1e-10 * float(g["svor_max"])
    finally:
        r.close()


@pytest.mark.gpu
def test_cuda_operators_match_golden():
    import ps3d_b200
    lib = ps3d_b200.load()
    g = np.load(os.path.join(G, "operators_16x32x8.npz"))
    lib.init(16, 32, 8, g["lower"], g["extent"])
    lib.init_inversion("Hou & Li")
    try:
        f = g["f"]
        for name, fn in (("fftxyp2s", lib.fftxyp2s), ("fftsine", lib.fftsine), ("fftcosine", lib.fftcosine),
                         ("diffx", lib.diffx), ("diffy", lib.diffy), ("diffz", lib.central_diffz),
                         ("combine", lib.field_combine_semi_spectral), ("decompose", lib.field_decompose_semi_spectral)):
            ref = g[name]
            assert np.max(np.abs(fn(f) - ref)) <= 1e-12 * max(np.max(np.abs(ref)), 1e-300), name
    finally:
        lib.finalise()


@pytest.mark.gpu
@pytest.mark.parametrize("stepper,tag", [("cn2", "cn2"), ("impl-diff-rk4", "rk4")])
def test_cuda_trajectory_matches_golden(stepper, tag):
    """100 steps of Beltrami 32^3: dt and adapt diagnostics every step, KE / enstrophy / helicity at the end to 1e-10."""
    import ps3d_b200
    from ps3d_b200 import host
    lib = ps3d_b200.load()
    g = np.load(os.path.join(G, f"beltrami32_{tag}_100steps.npz"))
    s = host.beltrami_solver(lib, 32, stepper=stepper)
    try:
        for i in range(100):
            dt, diag = s.advance()
            row = dict(zip(COLS, g["series"][i]))
            assert s.t == pytest.approx(row["t"], rel=1e-11) and dt == pytest.approx(row["dt"], rel=1e-11), i
            for k in COLS[2:]:
                assert diag[k] == pytest.approx(row[k], rel=1e-10), (i, k)
        lib.vor2vel()
        d = lib.diagnostics()
        for v, r in zip((d["ke"], d["en"], d["helicity"]), g["final"]):
            assert v == pytest.approx(r, rel=1e-10)
        svor = lib.download3("svor")[:, ::4, ::4, ::4]
        # accumulated over 100 steps, relative to the max-norm of the whole field (the sample misses the energetic modes)
        assert np.max(np.abs(svor - g["svor_sample"])) < 1e-10 * float(g["svor_max"])
    finally:
        s.close()


def test_gpu_series_512_matches_cpu_restatement_series():
    """BASELINE config 4 at FULL size: tests/golden/beltrami512_cn2_series.json was generated by the CUDA path on one
    B200 (tools/make_golden_512.py; bench.py checks every run against it), tests/golden/beltrami512_cn2_cpu_ref.json by
    the C++ restatement of the reference with the literal stafft kernels on the host cores of the same kind of box
    (tools/make_golden_512_cpu.py, ~2 min per step).  The two independent computations of the first steps of the
    benchmark trajectory must agree within the north_star bar: dt 1e-11, KE / enstrophy / helicity 1e-10."""
    import json
    gpu = json.load(open(os.path.join(G, "beltrami512_cn2_series.json")))
    cpu = json.load(open(os.path.join(G, "beltrami512_cn2_cpu_ref.json")))
    n = len(cpu["dt"])
    assert n >= 2 and cpu["grid"] == gpu["grid"] == 512
    for i in range(n):
        assert cpu["dt"][i] == pytest.approx(gpu["dt"][i], rel=1e-11), i
        for k in ("ke", "en", "helicity"):
            assert cpu[k][i] == pytest.approx(gpu[k][i], rel=1e-10), (k, i)
