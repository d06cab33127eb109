"""Dev harness: run the emulated kernels against the oracle (CPU only)."""
import sys, os, math, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ps3d_oracle as O
from ps3d_b200.lib import PS3DLib

EMU = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_emu", "libps3d_emu.so")

def err(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)

def main(nx=16, ny=32, nz=16):
    lib = PS3DLib(EMU)
    lower = np.array([-0.5 * math.pi] * 3); extent = np.array([math.pi] * 3)
    lib.init(nx, ny, nz, lower, extent)
    lib.init_inversion("Hou & Li")
    s = O.PS3D(nx, ny, nz, lower, extent)
    rng = np.random.default_rng(1234)
    f = rng.uniform(-1, 1, (nx, ny, nz + 1))
    fs = lib.fftxyp2s(f); print("fftxyp2s", err(fs, s.fftxyp2s(f)))
    print("fftxys2p", err(lib.fftxys2p(fs), f))
    print("fftsine", err(lib.fftsine(f), s.fftsine(f)))
    print("fftcosine", err(lib.fftcosine(f), s.fftcosine(f)))
    print("diffx", err(lib.diffx(f), s.diffx(f)))
    print("diffy", err(lib.diffy(f), s.diffy(f)))
    print("diffz", err(lib.central_diffz(f), s.central_diffz(f)))
    print("combine", err(lib.field_combine_semi_spectral(f), s.field_combine_semi_spectral(f)))
    print("decompose", err(lib.field_decompose_semi_spectral(f), s.field_decompose_semi_spectral(f)))
    print("combine_phys", err(lib.field_combine_physical(f), s.field_combine_physical(f)))
    print("decompose_phys", err(lib.field_decompose_physical(f), s.field_decompose_physical(f)))
    vor = rng.uniform(-1, 1, (3, nx, ny, nz + 1))
    ke_en = s.set_vorticity(vor)
    lib.upload_vorticity(vor)
    lib.vor2vel()
    for name in ("svor", "vor", "svel", "vel"):
        print("vor2vel", name, err(lib.download3(name), getattr(s, name)))
    d = lib.diagnostics(); print("diag", d, s.get_kinetic_energy(), s.get_enstrophy(), s.get_helicity())
    lib.source(); s.source()
    print("source svorts", err(lib.download3("svorts"), s.svorts))
    lib.init_diffusion(d["ke"], d["en"])
    for stepper in ("cn2", "impl-diff-rk4"):
        lib.stepper_setup(stepper)
        t = 0.0; to = 0.0
        for i in range(2):
            t, dt, diag = lib.advance(t, 100.0)
            to, dto = s.advance(to, 100.0, stepper, literal=True)
            print(stepper, i, "t", t, to, "dt", dt, dto, "svor", err(lib.download3("svor"), s.svor))
            print("   diag", {k: (diag[k], s.diag.get(k)) for k in ("vortmax", "vortrms", "vorch", "ggmax", "umax", "usggmax", "lsggmax")})
    lib.finalise()

if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:]] or [16, 32, 16]
    t0 = time.time(); main(*a); print("wall", time.time() - t0)
