"""CPU: the C-ABI shared library loads and exports every symbol include/tpat.h declares; argument
validation works without a GPU (no compute call is made here)."""
import ctypes
import os
import re

import conftest  # noqa: F401

HEADER = os.path.join(conftest.ROOT, "include", "tpat.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tpat_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from tpat import _lib
    names = declared_symbols()
    assert len(names) >= 14
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"libtpat.so does not export {n}"
    assert set(names) == set(_lib.SIGNATURES), "tpat/_lib.py SIGNATURES out of sync with include/tpat.h"


def test_version_and_struct_layout():
    from tpat import _lib
    assert _lib.lib.tpat_version() == _lib.TPAT_VERSION
    assert _lib.lib.tpat_sizeof_forward_args() == ctypes.sizeof(_lib.ForwardArgs)


def test_argument_validation_sets_error_message():
    from tpat import _lib
    lib = _lib.lib
    assert lib.tpat_layernorm(None, None, None, None, 0, 4, 768, 1e-6, None) != 0
    assert "null" in _lib.last_error()
    rc = lib.tpat_gemm(1 << 12, 0, 64, 1 << 12, 1, None, 1 << 12, 0, 64, None, 0, None, 0, 0, 8, 64, 64, 0, 0, None)
    assert rc != 0 and "dtype" in _lib.last_error()
    rc = lib.tpat_attention(1 << 12, 1 << 12, 1, None, 0, 2, 65, 12, 80, 1, 0.125, 1, None)
    assert rc != 0 and "head dim" in _lib.last_error()
    rc = lib.tpat_score_topk(1 << 12, 12, 12.0, None, 1 << 12, None, 2, 66, 2, 65, None)
    assert rc != 0 and "out of range" in _lib.last_error()
    args = _lib.ForwardArgs()
    assert lib.tpat_forward_workspace_bytes(ctypes.byref(args)) == 0
    assert lib.tpat_forward(ctypes.byref(args), None) != 0


def test_workspace_size_is_computed_on_cpu():
    from tpat import _lib
    a = _lib.ForwardArgs()
    a.variant, a.impl, a.B, a.T, a.F = 0, 1, 64, 1024, 128
    a.depth, a.D, a.H, a.Dh, a.num_classes = 12, 768, 12, 3072, 527
    for i in range(12):
        a.keep[i] = 512
    need = _lib.lib.tpat_forward_workspace_bytes(ctypes.byref(a))
    # x(2 fp32) + y + qkv(3) + ao (bf16) + hidden(4x bf16) dominate: ~ B*513*768*(8+2+6+2+8) bytes
    assert 0.6e9 < need < 0.8e9
    assert _lib.lib.tpat_forward_launch_count(ctypes.byref(a)) == 2 + 12 * 7 + 2
