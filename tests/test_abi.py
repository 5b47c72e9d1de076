"""CPU checks of the drop-in boundary: the library loads, exports every symbol include/gmc.h declares, the ctypes
table matches the header, and the product path fails loudly (no CPU fallback) without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_prototypes():
    txt = open(os.path.join(ROOT, "include", "gmc.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    protos = {}
    for m in re.finditer(r"GMC_API\s+([\w\s\*]+?)\s*\b(gmc_\w+)\s*\(([^;]*?)\)\s*;", txt, flags=re.S):
        args = [a.strip() for a in m.group(3).replace("\n", " ").split(",")]
        protos[m.group(2)] = [] if args == ["void"] else args
    return protos


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from mcmc_gpu_b200 import _lib
    return _lib.load()


def test_every_declared_symbol_is_exported_and_bound(lib):
    from mcmc_gpu_b200 import _lib
    protos = _header_prototypes()
    assert len(protos) >= 17
    for name, args in protos.items():
        assert hasattr(lib, name), f"{name} declared in gmc.h but not exported by libgmc.so"
        assert name in _lib.PROTOTYPES, f"{name} has no ctypes prototype"
        assert len(_lib.PROTOTYPES[name][1]) == len(args), f"{name}: ctypes arity differs from the header"
    assert set(_lib.PROTOTYPES) == set(protos)


def test_no_torch_types_in_the_abi():
    txt = open(os.path.join(ROOT, "include", "gmc.h")).read()
    assert "torch" not in re.sub(r"/\*.*?\*/", "", txt, flags=re.S) and "at::" not in txt


def test_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = lib.gmc_create(ctypes.byref(h), 0, 16, 16, 1)
    assert rc == -3 and b"no CPU fallback" in lib.gmc_last_error()
    from mcmc_gpu_b200 import Topography
    from mcmc_gpu_b200._lib import GmcError
    z = np.zeros((4, 4))
    with pytest.raises(GmcError):
        Topography.get_mass_conservation_residual(z, z, z, z, z, z, 1.0)


def test_null_and_bad_arguments(lib):
    assert lib.gmc_create(None, 0, 16, 16, 1) == -1
    h = ctypes.c_void_p()
    assert lib.gmc_create(ctypes.byref(h), 0, 1, 16, 1) == -2
    assert lib.gmc_residual(None, None, None, 1, None) == -1
    assert lib.gmc_destroy(None) == 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mcmc_gpu_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), f"{f} mentions the oracle"
