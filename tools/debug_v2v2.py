import sys, os, math
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ps3d_oracle as O
from ps3d_b200.lib import PS3DLib, LIB_PATH
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
path = sys.argv[2] if len(sys.argv) > 2 else LIB_PATH
lib = PS3DLib(path)
lower = -0.5 * math.pi * np.ones(3); extent = math.pi * np.ones(3)
lib.init(n, n, n, lower, extent); lib.init_inversion("Hou & Li")
s = O.PS3D(n, n, n, lower, extent)
vor = np.random.default_rng(7).uniform(-1, 1, (3, n, n, n + 1))
s.set_vorticity(vor)
# oracle intermediates (inversion.f90:86-139)
nz = n
svor = s.svor
ds = s.diffy(svor[0]) - s.diffx(svor[1])
d0 = ds[..., 0:1].copy(); dn = ds[..., nz:nz + 1].copy()
es0 = d0 * s.dthetam + dn * s.dthetap
dsg = ds.copy(); dsg[..., 1:nz] = s.green[..., 1:nz] * ds[..., 1:nz]
as_ = np.zeros_like(ds); as_[..., 1:nz] = s.rkz[1:nz] * dsg[..., 1:nz]
as_ = O.dct(as_, nz)
es = es0 + as_
dsd = dsg.copy(); dsd[..., 1:] = O.dst(dsg[..., 1:], nz)
def show(name, d, r):
    e = np.abs(d - r); m = np.abs(r).max()
    bad = np.argwhere(e > 1e-10 * m)
    print(name, e.max() / m, "nbad", len(bad), "z", sorted(set(bad[:, 2]))[:10] if len(bad) else [])
os.environ["PS3D_DBG"] = "1"
lib.upload_vorticity(vor); lib.vor2vel()
show("es", lib.download("svel", 0), es)
show("as", lib.download("svel", 1), as_)
show("dthm", lib.download("svel", 2), s.dthetam)
os.environ["PS3D_DBG"] = "2"
lib.upload_vorticity(vor); lib.vor2vel()
show("d0", lib.download("svel", 0), d0 + 0 * ds)
show("dn", lib.download("svel", 1), dn + 0 * ds)
show("dsd", lib.download("svel", 2)[..., 1:nz], dsd[..., 1:nz])
lib.finalise()
lib = PS3DLib(path)
lib.init(n, n, n, lower, extent); lib.init_inversion("Hou & Li")
os.environ["PS3D_DBG"] = "4"
lib.upload_vorticity(vor); lib.vor2vel()
t1 = lib.download("svel", 0); t2 = lib.download("svel", 1); t3 = lib.download("svel", 2)
for kx, ky in ((1, 1), (2, 3), (1, 0), (5, 7)):
    for zz in (0, 1, nz - 1, nz):
        print(kx, ky, zz, "dthm", t1[kx, ky, zz], s.dthetam[kx, ky, zz], "dthp", t2[kx, ky, zz], s.dthetap[kx, ky, zz], "thm", t3[kx, ky, zz], s.thetam[kx, ky, zz], "thp", t3[n - kx, ky, zz], s.thetap[kx, ky, zz])
lib.finalise()
def showbad(name, d, r):
    e = np.abs(d - r); m = np.abs(r).max()
    bad = np.argwhere(e > 1e-10 * m)
    print(name, e.max() / m, "nbad", len(bad), [tuple(int(v) for v in b) for b in bad[:60]])
showbad("dthm4", t1, s.dthetam)
showbad("dthp4", t2, s.dthetap)
