"""The C-ABI library loads, exports every symbol include/ps3d_cuda.h declares, and has no CPU path."""
import ctypes
import os
import re

import numpy as np
import pytest

import __graft_entry__ as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cuda_lib_path():
    return G.build_cuda()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ps3d_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ps3d_cuda_\w+)\s*\(", src)))


def test_header_and_binding_agree():
    from ps3d_b200.lib import EXPORTED_SYMBOLS
    assert declared_symbols() == EXPORTED_SYMBOLS


def test_library_exports_every_declared_symbol(cuda_lib_path):
    dll = ctypes.CDLL(cuda_lib_path)
    for name in declared_symbols():
        assert hasattr(dll, name), name


def test_no_cpu_fallback(cuda_lib_path):
    """Without a CUDA device init must fail loudly (PS3D_ERR_NO_DEVICE), not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from ps3d_b200.lib import PS3DLib, PS3DError
    lib = PS3DLib(cuda_lib_path)
    with pytest.raises(PS3DError) as e:
        lib.init(32, 32, 32, np.zeros(3), np.ones(3))
    assert e.value.status == 4
    with pytest.raises(PS3DError) as e:
        lib.vor2vel()
    assert e.value.status == 1          # not initialised


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ps3d_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().lower() or f in (), (dirpath, f)
