"""Config 2 of BASELINE.json: isolated transform + vor2vel microbenchmarks at 64^3 on one B200, white-noise field
numpy.random.default_rng(1234).uniform(-1, 1, (64, 64, 65)).  Resident kernels are timed with CUDA
events inside the library (ps3d_cuda_time_kernel); operator-mode calls (host buffers in and out) by wall clock.
Prints one JSON object.  Run on a GPU box: python tools/microbench_64.py [n]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ps3d_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lib = ps3d_b200.load()
lower = -0.5 * np.pi * np.ones(3)
extent = np.pi * np.ones(3)
lib.init(n, n, n, lower, extent)
lib.init_inversion("Hou & Li")
rng = np.random.default_rng(1234)
f = rng.uniform(-1, 1, (n, n, n + 1))
lib.upload_vorticity(rng.uniform(-1, 1, (3, n, n, n + 1)))
lib.vor2vel()
N = n * n * (n + 1)
out = {"grid": [n, n, n], "resident_kernels": {}, "operator_mode_ms": {}}
names = ["fwd_y_sweep", "fwd_x_sweep", "inv_x_sweep", "inv_y_sweep", "vor2vel_columns", "source_columns"]
alg = [16, 16, 16, 16, 8 * 16, 5 * 16]
for w, name in enumerate(names):
    lib.time_kernel(w, 5)
    ms = lib.time_kernel(w, 50)
    out["resident_kernels"][name] = {"ms": ms, "alg_GBs": (alg[w] * N / (ms * 1e-3) / 1e9) if ms > 0 else None}


def wall(fn, reps=20):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


fs = lib.fftxyp2s(f)
out["operator_mode_ms"]["fftxyp2s"] = wall(lambda: lib.fftxyp2s(f))
out["operator_mode_ms"]["fftxys2p"] = wall(lambda: lib.fftxys2p(fs))
out["operator_mode_ms"]["diffx"] = wall(lambda: lib.diffx(fs))
out["operator_mode_ms"]["diffy"] = wall(lambda: lib.diffy(fs))
out["operator_mode_ms"]["fftsine"] = wall(lambda: lib.fftsine(fs))
out["operator_mode_ms"]["fftcosine"] = wall(lambda: lib.fftcosine(fs))
out["operator_mode_ms"]["vor2vel (resident)"] = wall(lambda: lib.vor2vel())
out["roundtrip_error"] = float(np.max(np.abs(lib.fftxys2p(fs) - f)))
lib.finalise()
print(json.dumps(out))
