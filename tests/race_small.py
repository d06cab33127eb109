"""Test harness (dev): small vor2vel + source + one cn2 step per nz template, meant to run under
`compute-sanitizer --tool racecheck` (shared-memory hazards of the in-place column transforms) and
`--tool memcheck`; checks the result against the oracle as well."""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ps3d_oracle as O  # noqa: E402
from ps3d_b200.lib import PS3DLib, LIB_PATH  # noqa: E402

lib = PS3DLib(os.environ.get("PS3D_PROBE_LIB", LIB_PATH))
PI = math.pi
worst = 0.0
for nx, ny, nz in [(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(8, 8, 512), (8, 8, 64), (16, 16, 16)]:
    lo = np.array([-0.5 * PI, 0.0, -1.0]); ex = np.array([PI, 2 * PI, 2.0])
    lib.init(nx, ny, nz, lo, ex)
    lib.init_inversion("Hou & Li")
    s = O.PS3D(nx, ny, nz, lo, ex, "Hou & Li")
    vor = np.random.default_rng(11).uniform(-1, 1, (3, nx, ny, nz + 1))
    s.set_vorticity(vor)
    lib.upload_vorticity(vor)
    lib.vor2vel()
    errs = [np.max(np.abs(lib.download3(n) - getattr(s, n))) / np.max(np.abs(getattr(s, n))) for n in ("svor", "vor", "svel", "vel")]
    lib.source(); s.source()
    errs.append(np.max(np.abs(lib.download3("svorts") - s.svorts)) / np.max(np.abs(s.svorts)))
    # the fused cn2 update of the source kernel (ps3d_cuda_advance) and the spectral diffz of the buoyancy build
    d = lib.diagnostics()
    lib.init_diffusion(d["ke"], d["en"])
    s.init_diffusion(s.get_kinetic_energy(), s.get_enstrophy())
    lib.stepper_setup("cn2")
    lib.advance(0.0, 100.0); s.advance(0.0, 100.0, "cn2", literal=True)
    errs.append(np.max(np.abs(lib.download3("svor") - s.svor)) / np.max(np.abs(s.svor)))
    if (nz & (nz - 1)) == 0:
        fs = s.field_decompose_physical(np.random.default_rng(3).uniform(-1, 1, (nx, ny, nz + 1)))
        errs.append(np.max(np.abs(lib.diffz(fs) - s.diffz(fs))) / np.max(np.abs(s.diffz(fs))))
    print((nx, ny, nz), ["%.1e" % e for e in errs], flush=True)
    worst = max(worst, max(errs))
    lib.finalise()
assert worst < 1e-12, worst
print("race_small ok")
