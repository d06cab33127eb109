"""Dev tool: per-kernel device times for one (nx, ny, nz) grid: python tools/gpu_probe_shape.py nx ny nz"""
import json, math, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ps3d_b200 import host
from ps3d_b200.lib import PS3DLib, LIB_PATH

nx, ny, nz = (int(v) for v in sys.argv[1:4])
lib = PS3DLib(os.environ.get("PS3D_PROBE_LIB", LIB_PATH))
lower = -0.5 * math.pi * np.ones(3); extent = math.pi * np.ones(3)
s = host.Solver(lib, nx, ny, nz, lower, extent)
s.setup_fields(host.beltrami_vorticity(nx, ny, nz, lower, extent))
for _ in range(2):
    s.advance()
ms = []
for _ in range(3):
    s.advance(); ms.append(lib.last_advance_ms())
N = nx * ny * (nz + 1)
print(json.dumps(dict(grid=[nx, ny, nz], ms=ms, pts_per_s=nx * ny * nz / (min(ms) * 1e-3))))
names = ["fwd_y", "fwd_x", "inv_x", "inv_y", "vor2vel_spec", "source_spec"]
alg = [16, 16, 16, 16, 8 * 16, 5 * 16]
for w in range(6):
    lib.time_kernel(w, 2)
    t = lib.time_kernel(w, 10)
    print(f"  {names[w]:14s} {t:8.3f} ms  alg {alg[w] * N / (t * 1e-3) / 1e9:8.1f} GB/s")
s.close()
