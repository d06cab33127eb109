"""Test harness (dev): every kernel family on the CPU block emulator built with -fsanitize=address, to catch
shared-memory / global-memory overruns of the kernel index arithmetic (the emulator's shared memory and device
buffers are heap allocations).  Usage:
  g++ -O1 -g -std=c++17 -fopenmp -fPIC -shared -fsanitize=address -DPS3D_EMU -DPS3D_EMU_IMPL -x c++ \
      ps3d_b200/csrc/ps3d_cuda.cu -o /tmp/libps3d_emu_asan.so
  ASAN_OPTIONS=detect_leaks=0 LD_PRELOAD=$(gcc -print-file-name=libasan.so) python tests/asan_sweep.py /tmp/libps3d_emu_asan.so
"""
import sys, os, math, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ps3d_oracle as O
from ps3d_b200.lib import PS3DLib
lib=PS3DLib(sys.argv[1])
PI=math.pi
def rel(a,b): return np.max(np.abs(a-b))/max(np.max(np.abs(b)),1e-300)
for shape in [(8,8,8),(8,8,16),(8,8,32),(16,8,64),(8,8,128),(8,8,256),(8,8,512),(8,8,1024),(12,10,6),(10,12,30),(16,8,100)]:
    nx,ny,nz=shape
    lo=np.array([-0.5*PI,0.0,-1.0]); ex=np.array([PI,2*PI,2.0])
    lib.init(nx,ny,nz,lo,ex); lib.init_inversion("Hou & Li")
    s=O.PS3D(nx,ny,nz,lo,ex,"Hou & Li")
    rng=np.random.default_rng(11)
    f=rng.uniform(-1,1,(nx,ny,nz+1))
    lib.fftxyp2s(f); lib.fftsine(f); lib.fftcosine(f); lib.field_combine_physical(f); lib.field_decompose_physical(f); lib.central_diffz(f)
    vor=rng.uniform(-1,1,(3,nx,ny,nz+1))
    s.set_vorticity(vor); lib.upload_vorticity(vor); lib.vor2vel()
    d=lib.diagnostics(); lib.init_diffusion(d["ke"],d["en"]); lib.stepper_setup("cn2")
    t,dt,_=lib.advance(0.0,100.0); to,dto=s.advance(0.0,100.0,"cn2",literal=True)
    lib.vor2vel(); lib.adapt(t,100.0); lib.field_stats(); lib.genspec(); lib.pressure(); s.vor2vel()
    # buoyancy build (spectral diffz, tendency, bfmax, sbuoy updates) and an impl-diff-rk4 step (fused substeps)
    lib.enable_buoyancy(); lib.upload_buoyancy(rng.uniform(-0.5,0.5,(nx,ny,nz+1))); lib.set_physics((0.0,0.1,0.2),0.5)
    lib.init_diffusion_buoyancy(d["ke"],d["en"],3,20.0,"Kolmogorov","roll-mean-bfmax",2)
    lib.diffz(f); lib.advance(t,100.0); lib.stepper_setup("impl-diff-rk4"); lib.advance(t,100.0); lib.pressure()
    lib.upload_vorticity_begin(np.ascontiguousarray(vor)); lib.upload_vorticity_end()
    print(shape, "ok", flush=True)
    lib.finalise()
print("asan sweep done")
