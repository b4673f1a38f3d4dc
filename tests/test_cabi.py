"""The drop-in boundary: libb200timm.so loads and exports exactly the symbols include/b200timm.h declares
(no compute calls -- this runs without a GPU)."""
import ctypes
import os
import subprocess

import pytest

from b200timm import _lib


def test_header_and_binding_agree():
    hdr = _lib.header_symbols()
    assert hdr, "no B200_API prototypes found in include/b200timm.h"
    assert sorted(_lib.SIGNATURES) == hdr


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        pytest.fail(f"{_lib.LIB_PATH} missing: run __graft_entry__.build()")
    lib = _lib.load()
    for name in _lib.header_symbols():
        assert hasattr(lib, name), f"{name} not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l and "b200_" in l)
    assert exported == _lib.header_symbols(), "exports and header differ"


def test_abi_version_and_error_string_without_gpu():
    lib = _lib.load()
    assert lib.b200_abi_version() == 5
    assert isinstance(lib.b200_last_error(), bytes)


def test_gemm_desc_layout_matches_header():
    """struct b200_gemm_desc: 8 pointers, 5 int64, 12 int32."""
    assert ctypes.sizeof(_lib.GemmDesc) == 8 * 8 + 5 * 8 + 12 * 4 + 8 + 4 * 4
    assert _lib.GemmDesc.M.offset == 104 and _lib.GemmDesc.split_k.offset == 148 and _lib.GemmDesc.rope.offset == 152


def test_argument_errors_are_reported_not_crashed():
    lib = _lib.load()
    rc = lib.b200_gemm(None, None)
    assert rc == _lib.ERR_INVALID_ARG and b"null descriptor" in lib.b200_last_error()
    rc = lib.b200_attn_fwd(1, 1, None, None, 1, 16, 1, 32, 1.0, None)  # head_dim 32: unsupported, no fallback
    assert rc == _lib.ERR_UNSUPPORTED
    with pytest.raises(_lib.B200Error):
        _lib.check(rc, "b200_attn_fwd")
