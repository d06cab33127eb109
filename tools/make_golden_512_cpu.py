"""Runs the CPU restatement of the reference (oracle/ps3d_ref.cpp + the literal stafft kernels, all host cores) on the
FULL benchmark grid, Beltrami 512^3 cn2, for a few steps and records the dt sequence and KE / enstrophy / helicity at
the start of every step: an independent series for the first steps of tests/golden/beltrami512_cn2_series.json (which
the GPU path generated).  tests/test_golden.py compares the two committed files.  ~2 min per step on 16 cores.
usage: python tools/make_golden_512_cpu.py [n] [nsteps] [out]"""
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
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
out = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/beltrami%d_cn2_cpu_ref.json" % n
lower, extent = -0.5 * math.pi * np.ones(3), math.pi * np.ones(3)
r = RefSolver(n, n, n, lower, extent, path=G.build_ref(native=True))
r.set_vorticity(host.beltrami_vorticity(n, n, n, lower, extent))


def diagnostics():
    """field_diagnostics.f90:85-206 (trapezoid in z) of the vel / vor the last vor2vel left behind."""
    vel, vor = r.get("vel"), r.get("vor")
    w = np.ones(n + 1)
    w[0] = w[n] = 0.5
    ncelli = 1.0 / float(n) ** 3
    ke = 0.5 * float(np.einsum("cxyz,cxyz,z->", vel, vel, w)) * ncelli
    en = 0.5 * float(np.einsum("cxyz,cxyz,z->", vor, vor, w)) * ncelli
    he = float(np.einsum("cxyz,cxyz,z->", vel, vor, w)) * ncelli
    return ke, en, he


series = {"grid": n, "stepper": "cn2", "source": "oracle/ps3d_ref.cpp (C++/OpenMP restatement of the reference, literal stafft kernels)",
          "dt": [], "ke": [], "en": [], "helicity": [], "s_per_step": []}
for i in range(nsteps):
    t0 = time.perf_counter()
    t, dt = r.advance(stepper="cn2")
    series["s_per_step"].append(time.perf_counter() - t0)
    ke, en, he = diagnostics()          # of the vel / vor at the start of the step just taken
    series["dt"].append(dt); series["ke"].append(ke); series["en"].append(en); series["helicity"].append(he)
    print(i, dt, ke, en, he, flush=True)
r.close()
json.dump(series, open(out, "w"), indent=1)
g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "beltrami%d_cn2_series.json" % n)
if os.path.exists(g):
    gg = json.load(open(g))
    for k in ("dt", "ke", "en", "helicity"):
        print(k, [abs(a - b) / abs(b) for a, b in zip(series[k], gg[k])])
