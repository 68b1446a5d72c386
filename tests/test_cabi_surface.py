"""CPU: the C-ABI shared library loads and exports every symbol include/amcl3d_cuda.h declares; without a GPU
every compute entry fails loudly (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "amcl3d_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(amcl3d_cuda_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from amcl3d_b200 import capi
    syms = declared_symbols()
    assert len(syms) >= 35
    assert sorted(capi.SIGNATURES) == syms


def test_library_exports_every_declared_symbol():
    from amcl3d_b200 import capi
    lib = capi.load_library()
    for s in declared_symbols():
        assert hasattr(lib, s), s
    assert lib.amcl3d_cuda_abi_version() == 1


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under amcl3d_b200/ may reference it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "amcl3d_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"\boracle[./]|import oracle|amcl3d_oracle|libamcl3d_ref", src):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import amcl3d_b200
    with pytest.raises(amcl3d_b200.Amcl3dCudaError) as e:
        amcl3d_b200.Context(0)
    assert e.value.code == -1 and "no CPU fallback" in str(e.value)
