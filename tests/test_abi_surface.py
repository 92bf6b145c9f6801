"""The C-ABI library loads on a CPU-only host and exports every symbol that
include/b200arnoldi.h declares (no compute calls here)."""

import ctypes
import os
import re

import pytest

import b200arnoldi as b2a

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200arnoldi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2a_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_groups():
    syms = declared_symbols()
    for must in ("b2a_partialschur", "b2a_iterate_arnoldi", "b2a_orthogonalize", "b2a_reinitialize",
                 "b2a_rotate_basis", "b2a_rotate_final", "b2a_csr_create", "b2a_csc_create",
                 "b2a_op_from_callback", "b2a_ws_create", "b2a_basis_times"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(b2a.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_python_binding_covers_the_header():
    from arnoldimethod_jl_b200 import _lib

    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_version_and_error_string():
    lib = b2a.lib()
    assert lib.b2a_version() == 100
    assert isinstance(lib.b2a_last_error(), bytes)


def test_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(b2a.B200Error):
        b2a.Context(0)
