"""The drop-in boundary: libb200timm.so loads and exports exactly the symbols include/b200timm.h declares
(no compute calls -- this runs without a GPU)."""
import ctypes
import os
import subprocess

import pytest

from b200timm import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


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
    assert lib.b200_abi_version() == 8
    assert isinstance(lib.b200_last_error(), bytes)


def test_gemm_desc_layout_matches_header(tmp_path):
    """struct b200_gemm_desc as the C compiler lays it out from include/b200timm.h (gcc) == the ctypes mirror in _lib.py:
    total size and the offset of every field."""
    import subprocess
    fields = [f[0] for f in _lib.GemmDesc._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "b200timm.h"\nint main(void) {\n'
                   '  printf("%zu\\n", sizeof(b200_gemm_desc));\n'
                   + "".join(f'  printf("%zu\\n", offsetof(b200_gemm_desc, {f}));\n' for f in fields) + "  return 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    nums = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert ctypes.sizeof(_lib.GemmDesc) == nums[0]
    assert [getattr(_lib.GemmDesc, f).offset for f in fields] == nums[1:]
    assert fields[-1] == "ln_eps" and _lib.GemmDesc.M.offset == 104


def test_argument_errors_are_reported_not_crashed():
    lib = _lib.load()
    rc = lib.b200_gemm(None, None)
    assert rc == _lib.ERR_INVALID_ARG and b"null descriptor" in lib.b200_last_error()
    rc = lib.b200_attn_fwd(1, 1, None, None, 1, 16, 1, 32, 1.0, None)  # head_dim 32: unsupported, no fallback
    assert rc == _lib.ERR_UNSUPPORTED
    with pytest.raises(_lib.B200Error):
        _lib.check(rc, "b200_attn_fwd")
