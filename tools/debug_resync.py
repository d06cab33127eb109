import sys, os, math
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ps3d_oracle as O
from ps3d_b200.lib import PS3DLib, LIB_PATH
from ps3d_b200 import host
path = sys.argv[1] if len(sys.argv) > 1 else LIB_PATH
stepper = sys.argv[2] if len(sys.argv) > 2 else "cn2"
n = 32
lib = PS3DLib(path)
ref = O.beltrami_setup(n)
s = host.beltrami_solver(lib, n, stepper=stepper)
rng = np.random.default_rng(11)
ref.svor += 1e-3 * rng.uniform(-1, 1, ref.svor.shape) * ref.filt[None]
t = 0.0
def rel(a, b): return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)
for i in range(3):
    for c in range(3):
        lib.upload("svor", c, ref.svor[c])
    s.t = t
    dt, diag = s.advance()
    t, dto = ref.advance(t, 100.0, stepper, literal=True)
    d = lib.download3("svorts")
    e = np.abs(d - ref.svorts)
    print(i, "dt", dt, dto, "svor", rel(lib.download3("svor"), ref.svor), "svorts", rel(d, ref.svorts), "max|svorts|", np.abs(ref.svorts).max())
    bad = np.argwhere(e > 1e-9 * np.abs(ref.svorts).max())
    print("   nbad", len(bad), [tuple(int(v) for v in b) for b in bad[:12]])
s.close()
