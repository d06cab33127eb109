"""One-off: the C++/OpenMP restatement of the reference path (oracle/ps3d_ref.cpp + oracle/stafft_lit.c, -march=native)
on the FULL benchmark grid, Beltrami 512^3 cn2, on the GPU box's host cores: one warm-up step, then timed steps.
Writes profiles-ready JSON.  usage: python tools/cpu_512_once.py [n] [steps] [out]"""
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as G  # noqa: E402
from oracle.ps3d_ref import RefSolver  # noqa: E402
from ps3d_b200 import host  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
out = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/r02_cpu_%d.json" % n
lower, extent = -0.5 * math.pi * np.ones(3), math.pi * np.ones(3)
t0 = time.perf_counter()
r = RefSolver(n, n, n, lower, extent, path=G.build_ref(native=True))
r.set_vorticity(host.beltrami_vorticity(n, n, n, lower, extent))
setup = time.perf_counter() - t0
r.advance(stepper="cn2")
t0 = time.perf_counter()
for _ in range(steps):
    r.advance(stepper="cn2")
sec = (time.perf_counter() - t0) / steps
res = {"workload": "Beltrami %d^3 cn2 (examples/beltrami_512.config)" % n, "value": n ** 3 / sec, "unit": "grid-pt*steps/s",
       "s_per_step": sec, "steps": steps, "warmup": 1, "setup_s": setup, "cores": os.cpu_count(), "threads": r.threads,
       "kind": "port", "build": "g++ -O3 -march=native -fopenmp (oracle/ps3d_ref.cpp + oracle/stafft_lit.c)",
       "note": "restatement of the reference algorithm with its sweep structure and its own FFT kernels; not the Fortran build"}
r.close()
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res))
