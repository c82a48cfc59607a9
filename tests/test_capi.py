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
