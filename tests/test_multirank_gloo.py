"""N > 1 path on CPU: two ranks of the x-slab / ky-slab decomposition run the emulated kernels, exchanging
through torch.distributed (gloo) via ps3d_cuda_set_transport, and must reproduce the one-rank oracle."""
import os
import subprocess
import sys

import pytest

import __graft_entry__ as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("stepper,n", [("cn2", 16), ("impl-diff-rk4", 16), ("cn2", 12)])
def test_two_rank_slab_decomposition_matches_oracle(stepper, n):
    """n = 12: a grid that is not a power of two (slab blocks of 6 rows: dividing row maps of line_gen.cuh)."""
    emu = G.build_emu()
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "multirank_worker.py"), emu, stepper, str(n)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("worst") == 2


@pytest.mark.parametrize("stepper", ["cn2", "impl-diff-rk4"])
def test_two_rank_buoyancy_matches_oracle(stepper):
    """The ENABLE_BUOYANCY paths (buoyancy tendency, bfmax reduced over the ranks, sbuoy stepping, f_cor) on two ranks."""
    emu = G.build_emu()
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29613", os.path.join(ROOT, "tests", "multirank_worker.py"), emu, stepper, "16", "buoyancy"]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("worst") == 2
