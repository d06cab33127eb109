import sys, os, math
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ps3d_oracle as O
import ps3d_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lib = ps3d_b200.load()
lower = -0.5 * math.pi * np.ones(3); extent = math.pi * np.ones(3)
lib.init(n, n, n, lower, extent); lib.init_inversion("Hou & Li")
s = O.PS3D(n, n, n, lower, extent)
vor = np.random.default_rng(7).uniform(-1, 1, (3, n, n, n + 1))
s.set_vorticity(vor)
lib.upload_vorticity(vor)
lib.vor2vel()
for name in ("svor", "vor", "svel", "vel"):
    d = lib.download3(name); r = getattr(s, name)
    e = np.abs(d - r)
    print(name, e.max() / np.abs(r).max())
    if e.max() / np.abs(r).max() > 1e-10:
        for c in range(3):
            bad = np.argwhere(e[c] > 1e-10 * np.abs(r).max())
            print("  comp", c, "nbad", len(bad), "kx", sorted(set(bad[:, 0]))[:20], "ky", sorted(set(bad[:, 1]))[:20], "z", sorted(set(bad[:, 2]))[:20])
lib.source(); s.source()
print("svorts", np.abs(lib.download3("svorts") - s.svorts).max() / np.abs(s.svorts).max())
lib.finalise()
