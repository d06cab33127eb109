"""Generates tests/golden/beltrami512_cn2_series.json ON A GPU BOX (one GPU): the dt sequence and KE / enstrophy /
helicity of the first NSTEPS cn2 steps of Beltrami 512^3 (BASELINE config 4).  bench.py compares the timed steps of
every run (any number of GPUs) with it and prints the result as `parity` on its JSON line, so that the scaling runs
prove they compute the same trajectory.  The one-GPU path itself is checked against the oracle at 32^3..256^3 and
against the analytic known answers at 512^3 (tests/test_gpu_parity.py); KE and enstrophy at t = 0 are analytic.
usage: python tools/make_golden_512.py [n] [nsteps] [out]"""
import json
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ps3d_b200  # noqa: E402
from ps3d_b200 import host  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/beltrami%d_cn2_series.json" % n
lib = ps3d_b200.load()
s = host.beltrami_solver(lib, n, stepper="cn2")
d0 = lib.diagnostics()
series = {"grid": n, "stepper": "cn2", "config": "examples/beltrami_512.config (Hou & Li, nnu 3, prediss 30, vorch, alpha 0.1)",
          "initial": {k: float(d0[k]) for k in ("ke", "en", "helicity")}, "dt": [], "ke": [], "en": [], "helicity": []}
for i in range(nsteps):
    dt, diag = s.advance()
    d = lib.diagnostics()          # of vel / vor at the start of the step just taken (what write_step would see)
    series["dt"].append(float(dt))
    for k in ("ke", "en", "helicity"):
        series[k].append(float(d[k]))
s.close()
# analytic values at t = 0 for the Beltrami flow k = l = 2, m = 1 on [-pi/2, pi/2]^3: <|omega|^2> = alpha^2 <|u|^2>
series["analytic_en_over_ke"] = 9.0
json.dump(series, open(out, "w"), indent=1)
print(json.dumps({"out": out, "dt0": series["dt"][0], "ke0": series["initial"]["ke"], "en0/ke0": series["initial"]["en"] / series["initial"]["ke"]}))
