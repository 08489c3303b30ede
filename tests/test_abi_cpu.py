"""The C-ABI shared library loads without a GPU and exports every symbol include/*.h declares."""
import ctypes
import os
import re

import pytest

from tf2_gnn_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "tfgnn_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tfgnn_b200_\w+)\s*\(", text)))


def test_library_is_built_in_tree():
    assert os.path.exists(_ffi.library_path())
    assert os.path.dirname(_ffi.library_path()).endswith(os.path.join("tf2_gnn_b200", "csrc"))


def test_every_declared_symbol_is_exported():
    lib = ctypes.CDLL(_ffi.library_path())
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/tfgnn_b200.h but not exported"
    assert set(_ffi.EXPORTED_SYMBOLS) == set(declared)


def test_abi_version_and_error_string():
    lib = _ffi.lib()
    assert lib.tfgnn_b200_abi_version() == 1
    assert isinstance(lib.tfgnn_b200_last_error(), bytes)


def test_signatures_do_not_mention_torch():
    with open(os.path.join(ROOT, "include", "tfgnn_b200.h")) as f:
        text = f.read()
    assert "torch" not in text.lower() and "at::" not in text


def test_invalid_arguments_are_reported_without_a_gpu():
    """Argument validation happens before any CUDA call, so it is testable on CPU."""
    lib = _ffi.lib()
    out = ctypes.c_void_p()
    rc = lib.tfgnn_b200_prepare(None, None, 99, 10, 0, ctypes.byref(out), None)
    assert rc == _ffi.ERR_INVALID_ARGUMENT
    with pytest.raises(ValueError):
        _ffi.check(rc)
    assert b"num_edge_types" in lib.tfgnn_b200_last_error()
    rc = lib.tfgnn_b200_dense_fwd(None, None, None, 4, 0, 4, 0, 0, None)
    assert rc == _ffi.ERR_INVALID_ARGUMENT
