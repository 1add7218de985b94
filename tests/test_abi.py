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


def test_argument_validation_returns_error_codes(lib):
    """Bad arguments are rejected with a negative status and a message before anything is launched (so this needs
    no device): the reference surface raises Python exceptions, the shim turns these codes into LdxError."""
    l = lib.lib()
    buf = ctypes.create_string_buffer(256)
    addr = ctypes.addressof(buf)
    bad = [
        l.ldx_dpm_eps(None, addr, 1.0, addr, 4, None),
        l.ldx_dpm_eps(addr, addr, 0.0, addr, 4, None),                     # sigma must be positive
        l.ldx_dpm_combine(addr, None, None, 1.0, 0.0, addr, 4, None),
        l.ldx_dpm_error_norm(addr, addr, addr, 0.01, 0.05, 4, None, addr, None),
        l.ldx_bislerp_axis(addr, addr, addr, addr, addr, 1, 17, 1, 4, 8, 1, None),   # more than 16 channels
        l.ldx_bislerp_axis(addr, addr, None, addr, addr, 1, 4, 1, 4, 8, 1, None),
        l.ldx_clip_embed(addr, addr, addr, addr, 4, 2, 77, 768, None),     # row pitch below the width
        l.ldx_vae_preprocess(None, addr, 4, None),
        l.ldx_vae_sample_posterior(addr, None, None, 1, 4, 16, None),
        l.ldx_attention_causal(addr, 64, addr, 64, addr, 64, addr, 64, 1, 77, 1, 40, None),   # d = 64 only
        l.ldx_cfg_dpmpp2m_step(addr, 8, addr, None, None, 1.0, 7.0, 0.5, 0.5, 0.0, 0, 0, 0.5, None, 1, 4, 8, 8, None),
        l.ldx_cfg_dpmpp2m_step(addr, 2, addr, addr, None, 1.0, 7.0, 0.5, 0.5, 0.0, 0, 0, 0.5, None, 1, 4, 8, 8, None),  # pitch < C
        l.ldx_vae_postprocess_u8(addr, 8, None, 16, None),
        # fused feed-forward: built for C = 320 only; F a multiple of 64; no null operands
        l.ldx_ff_geglu_fused(addr, 640, 256, addr, 640, addr, addr, addr, 2560, addr, addr, 1, 1e-5, 640, 2560, addr, 640, None),
        l.ldx_ff_geglu_fused(addr, 320, 256, addr, 320, addr, addr, addr, 1280, addr, addr, 1, 1e-5, 320, 1250, addr, 320, None),
        l.ldx_ff_geglu_fused(addr, 320, 256, addr, 320, addr, None, addr, 1280, addr, addr, 1, 1e-5, 320, 1280, addr, 320, None),
    ]
    assert all(rc != 0 for rc in bad), bad
    assert l.ldx_last_error_string()
    with pytest.raises(lib.LdxError):
        lib.check(bad[0], "ldx_dpm_eps")


def test_halo_conv_descriptor_validation(lib):
    """The halo-tile conv's descriptor rules (include/ldx.h: halo_nsrc, gn_coef) are checked before any device work."""
    l = lib.lib()
    buf = ctypes.create_string_buffer(4096)
    addr = ctypes.addressof(buf)

    def desc(W=16, H=16, C=64, halo=1, flags=0, coef=None, nseg=0):
        d = lib.GemmDesc()
        d.nsrc = 1
        d.src[0].ptr, d.src[0].C, d.src[0].W, d.src[0].H, d.src[0].B = addr, C, W, H, 1
        d.src[0].sw, d.src[0].sh, d.src[0].sb = C, C * W, C * W * H
        d.nseg, d.w, d.ldw, d.N = nseg, addr, 9 * 64, 64
        d.W, d.H, d.B, d.out = W, H, 1, addr
        d.o_sx, d.o_sy, d.o_sb = 64, 64 * W, 64 * W * H
        d.flags, d.halo_nsrc = flags, halo
        if coef:
            d.gn_coef, d.gn_ctot = coef, C
        return d

    bad = [
        l.ldx_gemm(ctypes.byref(desc(W=12)), None),                         # W % 8 != 0
        l.ldx_gemm(ctypes.byref(desc(halo=3)), None),                       # at most two halo sources
        l.ldx_gemm(ctypes.byref(desc(flags=lib.EPI_OUT_F32)), None),        # fp16 output only
        l.ldx_gemm(ctypes.byref(desc(flags=lib.EPI_GEGLU)), None),
        l.ldx_gemm(ctypes.byref(desc(C=60)), None),                         # C % 8 != 0
        l.ldx_gemm(ctypes.byref(desc(coef=addr + 4)), None),                # coefficient table alignment
        l.ldx_gemm(ctypes.byref(desc(halo=0, nseg=0)), None),               # a plain GEMM needs at least one segment
        l.ldx_group_norm_coef(addr, 64, 64, None, 0, 0, 1, 256, 1e-5, addr, addr, addr, None, addr, None, 0, None, 0, None),
        l.ldx_group_norm_coef(addr, 64, 64, None, 0, 0, 1, 256, 1e-5, addr, addr, addr, addr, None, None, 0, None, 0, None),
    ]
    assert all(rc == -1 for rc in bad), bad                                 # LDX_E_INVALID, nothing launched


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under lightdiffusion_b200/ may import it, except
    smoke.py, which uses it only as the checker."""
    pkg = os.path.join(ROOT, "lightdiffusion_b200")
    pat = re.compile(r"^\s*(from\s+oracle|import\s+oracle)", re.M)
    for fn in os.listdir(pkg):
        if fn.endswith(".py") and fn != "smoke.py":
            assert not pat.search(open(os.path.join(pkg, fn)).read()), fn
