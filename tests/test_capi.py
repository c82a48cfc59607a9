"""The C-ABI library loads and exports every symbol include/igi_b200.h declares (no compute)."""
import ctypes

from isaacgyminsertion_b200 import _lib


def test_header_symbols_exported(built_lib):
    lib = ctypes.CDLL(built_lib)
    names = _lib.declared_symbols()
    assert "igi_pcl_compact" in names and "igi_fps" in names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/igi_b200.h but not exported"


def test_version_and_error_string(built_lib):
    lib = _lib.load()
    assert lib.igi_version() == 100
    assert isinstance(lib.igi_last_error(), bytes)


def test_bad_args_return_error_without_gpu(built_lib):
    lib = _lib.load()
    rc = lib.igi_fps(None, ctypes.c_int64(0), None, None, ctypes.c_int64(1), 0, 1, 4, None,
                     ctypes.c_int64(12), None, 0, None)
    assert rc == -1
    assert b"igi_fps" in lib.igi_last_error()


def test_cpu_tensor_is_rejected(built_lib):
    import pytest
    import torch
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        _lib.dptr(torch.zeros(4), torch.float32, "x")


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No CPU fallback: without the built .so the loader raises instead of routing anywhere else."""
    import pytest
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libigi_b200.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_engines_refuse_a_cpu_device(built_lib):
    import numpy as np
    import pytest
    from isaacgyminsertion_b200.allsight_render import BatchedAllSight
    with pytest.raises(RuntimeError):
        BatchedAllSight(2, np.zeros(2, dtype=np.int64), device="cpu")


def test_product_package_never_imports_the_oracle():
    import os
    import re
    pkg = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "isaacgyminsertion_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
