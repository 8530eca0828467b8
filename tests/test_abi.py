"""The C-ABI library: loads, exports every symbol include/tvf.h declares, refuses to run without a GPU
(no CPU fallback), and the host mirror reproduces the reference's error behaviour."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "tvf.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tvf_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(libtvf_path):
    lib = ctypes.CDLL(libtvf_path)
    names = _declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), "libtvf.so does not export %s" % name


def test_binding_matches_header(libtvf_path):
    from tft_vs_fund_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    lib = _lib.load()
    assert lib.tvf_version() >= 100


def test_no_cpu_fallback(libtvf_path):
    """Without a CUDA device handle creation must fail loudly (and with one, succeed)."""
    from tft_vs_fund_b200 import _lib
    lib = _lib.load()
    if lib.tvf_device_count() == 0:
        with pytest.raises(_lib.TvfError, match="no CUDA device"):
            _lib.Handle(0)
        import tft_vs_fund_b200 as pkg
        with pytest.raises(_lib.TvfError):
            pkg.LinearTFTPoseEstimation(np.zeros((6, 20)), np.zeros((9, 3)))
    else:
        _lib.Handle(0).close()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "tft_vs_fund_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "hostcheck" not in text or f.endswith(".cuh"), f


def test_linearF_error_before_any_device_work():
    import tft_vs_fund_b200 as pkg
    p = np.zeros((2, 7))
    with pytest.raises(ValueError, match="At least 8 correspondences are necessary"):   # linearF.m:35-37
        pkg.linearF(p, p)
    with pytest.raises(ValueError):
        pkg.linearF(np.zeros((2, 9)), np.zeros((2, 8)))
    with pytest.raises(ValueError, match="At least 8 correspondences are necessary"):   # optimF.m:36-38
        pkg.optimF(p, p)
    with pytest.raises(ValueError, match="At least 8 correspondences are necessary"):
        pkg.OptimFPoseEstimation(np.zeros((6, 7)), np.zeros((9, 3)))
