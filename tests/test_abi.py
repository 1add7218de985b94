"""CPU-side checks of the drop-in boundary: libldx.so builds, loads, exports every symbol that
include/ldx.h declares, the ctypes structs match the C structs, and compute calls fail loudly without a
GPU (there is no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from lightdiffusion_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib


def header_functions():
    src = open(os.path.join(ROOT, "include", "ldx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ldx_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(lib):
    names = header_functions()
    assert len(names) >= 25
    dll = ctypes.CDLL(lib.LIB_PATH)
    missing = [n for n in names if not hasattr(dll, n)]
    assert not missing, missing
    # and the binding covers the whole header
    unbound = [n for n in names if n not in lib.EXPORTS]
    assert not unbound, unbound


def test_struct_sizes_match(lib):
    l = lib.lib()          # raises on ABI mismatch
    assert l.ldx_sizeof(0) == ctypes.sizeof(lib.GemmDesc)
    assert l.ldx_sizeof(1) == ctypes.sizeof(lib.GenericOp)
    assert l.ldx_version() >= 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    l = lib.lib()
    assert l.ldx_device_check() != 0
    assert b"CUDA" in l.ldx_last_error_string() or b"device" in l.ldx_last_error_string()
    with pytest.raises(lib.LdxError):
        lib.require_device()
    from lightdiffusion_b200 import sd15
    from lightdiffusion_b200.engine import UNetEngine
    with pytest.raises(Exception):
        UNetEngine({}, sd15.tiny_unet_config())
    from lightdiffusion_b200 import ops
    conv = ops.ldx_ops.Conv2d(8, 8, 3, padding=1)
    with pytest.raises(lib.LdxError):
        conv(torch.zeros(1, 8, 4, 4))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under lightdiffusion_b200/ may import it, except
    smoke.py, which uses it only as the checker."""
    pkg = os.path.join(ROOT, "lightdiffusion_b200")
    pat = re.compile(r"^\s*(from\s+oracle|import\s+oracle)", re.M)
    for fn in os.listdir(pkg):
        if fn.endswith(".py") and fn != "smoke.py":
            assert not pat.search(open(os.path.join(pkg, fn)).read()), fn
