"""Quick timing probe (dev tool): per-kernel and per-step device times on one GPU."""
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ps3d_b200
from ps3d_b200 import host

from ps3d_b200.lib import PS3DLib, LIB_PATH
lib = PS3DLib(os.environ.get('PS3D_PROBE_LIB', LIB_PATH))
for n in [int(v) for v in sys.argv[1:]] or [256, 512]:
    for stepper in ("cn2",):
        s = host.beltrami_solver(lib, n, stepper=stepper)
        for _ in range(2):
            s.advance()
        ms = []
        for _ in range(5):
            s.advance(); ms.append(lib.last_advance_ms())
        N = n * n * (n + 1)
        sweeps = 115 if stepper == "cn2" else 150
        gbs = sweeps * 16 * N / (min(ms) * 1e-3) / 1e9
        print(json.dumps(dict(n=n, stepper=stepper, tma_launches=lib.tma_launches(), launches=lib.kernel_launches(), ms=ms, pts_per_s=n ** 3 / (min(ms) * 1e-3), alg_GBs=gbs)))
        if stepper == "cn2":
            names = ["fwd_y", "fwd_x", "inv_x", "inv_y", "vor2vel_spec", "source_spec"]
            alg = [16, 16, 16, 16, 8 * 16, 5 * 16]
            for w in range(6):
                lib.time_kernel(w, 2)
                t = lib.time_kernel(w, 10)
                print(f"  {names[w]:14s} {t:8.3f} ms  alg {alg[w] * N / (t * 1e-3) / 1e9:8.1f} GB/s")
        s.close()
