"""The C-ABI library loads and exports exactly the entry points include/pwknn.h declares
(no compute calls: this runs without a GPU), and the ctypes binding covers all of them."""

import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "pwknn.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pwknn_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported_and_bound():
    from pathway_b200 import _native

    if not os.path.exists(_native.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    lib = _native.lib()
    declared = _declared()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in pwknn.h but not exported by libpwknn.so"
        assert name in _native.SIGNATURES, f"{name} has no ctypes signature in pathway_b200/_native.py"
    out = subprocess.check_output(["nm", "-D", "--defined-only", _native.LIB_PATH], text=True)
    exported = sorted(set(re.findall(r" T (pwknn_[a-z0-9_]+)", out)))
    assert exported == declared
    assert lib.pwknn_abi_version() == 2


def test_no_gpu_fails_loudly():
    """Without a usable B200 the product refuses to create an index: there is no CPU fallback."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pathway_b200._native import PwknnError
    from pathway_b200.external_index import B200BruteForceKNNIndex

    with pytest.raises(PwknnError) as e:
        B200BruteForceKNNIndex(8, 4)
    assert e.value.code == 3  # PWKNN_E_NOGPU


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    for dirpath, _dirs, files in os.walk(os.path.join(ROOT, "pathway_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower(), (dirpath, f)
