"""The C-ABI library loads without a GPU and exports every symbol include/lia_ral_b200.h
declares; compute entry points fail loudly (no CPU fallback) when no CUDA device exists."""
import ctypes as ct
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from lia_ral_b200 import capi
    capi.build()
    return capi.lib()


def _declared():
    src = open(os.path.join(ROOT, "include", "lia_ral_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lr_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = _declared()
    assert len(names) > 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_version_and_error_plumbing(lib):
    assert b"sm_100a" in lib.lr_version()
    assert lib.lr_set_gmm_kernel(7) != 0
    assert b"selector" in lib.lr_last_error()
    assert lib.lr_set_gmm_kernel(0) == 0


def test_no_cpu_fallback_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from lia_ral_b200 import capi
    with pytest.raises(capi.LrError) as ei:
        capi.GMM(np.ones(2) / 2, np.zeros((2, 3)), np.ones((2, 3)))
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)
    with pytest.raises(capi.LrError):
        capi.init(0)
