"""Torch-tensor level wrappers over the libldx.so C ABI.

Activations are fp16 CUDA tensors in channels-last logical layout: a 4-D tensor is indexed
[B, H, W, C] with ``stride(-1) == 1`` (channel slices and strided spatial views are fine); a 2-D
tensor is [rows, C].  Every function enqueues on the current torch CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import torch

from . import _lib as L


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def ceil64(c: int) -> int:
    return (c + 63) // 64 * 64


def _check_act(t: torch.Tensor):
    if not (t.is_cuda and t.dtype == torch.float16 and t.stride(-1) == 1):
        raise L.LdxError(f"expected a CUDA fp16 tensor with unit channel stride, got {t.dtype} {t.device} {t.stride()}")


def make_src(t: torch.Tensor) -> L.Src:
    _check_act(t)
    s = L.Src()
    s.ptr = t.data_ptr()
    if t.dim() == 2:
        m, k = t.shape
        s.C, s.W, s.H, s.B = k, m, 1, 1
        s.sw = t.stride(0)
        s.sh = s.sb = t.stride(0) * max(m, 1)
    else:
        b, h, w, c = t.shape
        s.C, s.W, s.H, s.B = c, w, h, b
        s.sw, s.sh, s.sb = t.stride(2), t.stride(1), t.stride(0)
    return s


def _fill_out(d: L.GemmDesc, out: torch.Tensor, n_out: int):
    if out.dim() == 2:
        m = out.shape[0]
        d.W, d.H, d.B = m, 1, 1
        d.o_sx, d.o_sy, d.o_sb = out.stride(0), 0, 0
    else:
        b, h, w, _ = out.shape
        d.W, d.H, d.B = w, h, b
        d.o_sx, d.o_sy, d.o_sb = out.stride(2), out.stride(1), out.stride(0)
    d.o_off = 0
    d.out = out.data_ptr()
    assert out.shape[-1] >= n_out and out.stride(-1) == 1


def gemm_desc(srcs: Sequence[torch.Tensor], segs: Sequence[tuple], w: torch.Tensor, out: torch.Tensor, n: int,
              bias: Optional[torch.Tensor] = None, rowadd: Optional[torch.Tensor] = None,
              residual: Optional[torch.Tensor] = None, flags: int = 0, bn: int = 0,
              rownorm: Optional[tuple] = None) -> L.GemmDesc:
    """segs: (src_index, c0, dx, dy, nchunks).  w: fp16 [N, K] (K = 64 * total chunks, row stride free).
    rownorm = (stats [M, parts, 2] fp32, gsum [N] fp32, eps, C): LayerNorm of the A rows folded into the GEMM
    (w, bias must already be the folded w * gamma and bias + w @ beta, see fold_layer_norm)."""
    d = L.GemmDesc()
    d.nsrc = len(srcs)
    for i, t in enumerate(srcs):
        d.src[i] = make_src(t)
    d.nseg = len(segs)
    for i, (si, c0, dx, dy, nch) in enumerate(segs):
        d.seg[i] = L.Seg(si, c0, dx, dy, nch)
    assert w.dtype == torch.float16 and w.is_cuda and w.stride(1) == 1
    d.w, d.ldw, d.N = w.data_ptr(), w.stride(0), n
    n_out = n // 2 if flags & L.EPI_GEGLU else n
    _fill_out(d, out, n_out)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
        d.bias = bias.data_ptr()
    if rowadd is not None:
        assert rowadd.dtype == torch.float32 and rowadd.stride(1) == 1
        d.rowadd, d.ld_rowadd = rowadd.data_ptr(), rowadd.stride(0)
    if residual is not None:
        _check_act(residual)
        r2 = residual if residual.dim() == 2 else residual
        if residual.dim() == 4:
            b, h, w_, c = residual.shape
            assert residual.stride(1) == residual.stride(2) * w_ and residual.stride(0) == residual.stride(1) * h, \
                "residual must be pixel-contiguous"
            d.ldr = residual.stride(2)
        else:
            d.ldr = r2.stride(0)
        d.residual = residual.data_ptr()
    if rownorm is not None:
        stats, gsum, eps, c = rownorm
        assert stats.dtype == torch.float32 and stats.is_contiguous() and stats.dim() == 3 and stats.shape[2] == 2
        assert gsum.dtype == torch.float32 and gsum.is_contiguous() and gsum.numel() == n
        d.rowstats_in, d.rowstats_in_parts = stats.data_ptr(), stats.shape[1]
        d.ln_gsum, d.ln_eps, d.ln_c = gsum.data_ptr(), float(eps), int(c)
        flags |= L.EPI_ROWNORM
    d.flags, d.bn = flags, bn
    return d


def rowstats_parts(d: L.GemmDesc) -> int:
    """Column parts per row that GEMM `d` will write into its rowstats_out buffer."""
    n = L.lib().ldx_gemm_rowstats_parts(C.byref(d))
    if n < 0:
        raise L.LdxError("ldx_gemm_rowstats_parts: " + L.lib().ldx_last_error_string().decode())
    return n


def attach_rowstats(d: L.GemmDesc, rows: int) -> torch.Tensor:
    """Allocates [rows, parts, 2] fp32 and makes GEMM `d` write its per-row output statistics there."""
    stats = torch.empty(rows, rowstats_parts(d), 2, dtype=torch.float32, device="cuda")
    d.rowstats_out, d.rowstats_out_parts = stats.data_ptr(), stats.shape[1]
    return stats


def colstats_slots(d: L.GemmDesc) -> int:
    """32-row slots per batch element GEMM `d` would write into colstats_out; -1 when it cannot (split-K, direct-store
    epilogue, fewer than 32 pixels per batch element)."""
    return L.lib().ldx_gemm_colstats_slots(C.byref(d))


def attach_colstats(descs, batch: int, n: int):
    """Makes the GEMM(s) `descs` (one, or the four phase GEMMs of an upsample conv that together write one tensor)
    leave the per-(32-row slot, column) sums of their fp16 outputs in a shared [batch, slots, n, 2] fp32 buffer.
    Returns (buffer, slots) for group_norm_op(colstats=...), or None when a launch cannot produce them."""
    descs = descs if isinstance(descs, (list, tuple)) else [descs]
    per = [colstats_slots(d) for d in descs]
    if min(per) < 1:
        return None
    total = sum(per)
    buf = torch.empty(batch, total, n, 2, dtype=torch.float32, device="cuda")
    base = 0
    for d, s in zip(descs, per):
        d.colstats_out, d.colstats_slots, d.colstats_slot_base = buf.data_ptr(), total, base
        base += s
    return buf, total


def _dev16(t: torch.Tensor) -> torch.Tensor:
    """fp16 copy on the GPU; host tensors are converted on the host first (one H2D copy, no cast kernel)."""
    return t.detach().to(torch.float16).contiguous().to("cuda")


def fold_layer_norm(w: torch.Tensor, bias: Optional[torch.Tensor], gamma: torch.Tensor, beta: torch.Tensor):
    """LayerNorm(gamma, beta) followed by Linear(w, bias)  ->  (w' fp16 [N, K], gsum fp32 [N], bias' fp32 [N]) for
    LDX_EPI_ROWNORM: w' = w * gamma, gsum = rowsum(w'), bias' = bias + w @ beta (F.layer_norm LD.py:2403 + F.linear).
    A one-time load-time transform, computed where the weights live (host or device) with elementwise products and
    row sums only."""
    src = w.device
    w32 = w.detach().to(torch.float16).float()
    g32, b32 = gamma.detach().to(src, torch.float32), beta.detach().to(src, torch.float32)
    wf = (w32 * g32[None, :]).to(torch.float16)
    gsum = wf.float().sum(1)
    bf = (w32 * b32[None, :]).sum(1)
    if bias is not None:
        bf = bf + bias.detach().to(src, torch.float32)
    return wf.contiguous().to("cuda"), gsum.contiguous().to("cuda"), bf.contiguous().to("cuda")


# ----------------------------------------------------------------------------- descriptor builders
def linear_desc(a: torch.Tensor, w: torch.Tensor, out: torch.Tensor, **kw) -> L.GemmDesc:
    """out[M, N] = a[M, K] @ w[N, K]^T (+ epilogue).  K is taken from a; w may be K-padded to 64."""
    k = a.shape[-1]
    a2 = a if a.dim() == 2 else a.reshape(-1, k)
    o2 = out if out.dim() == 2 else out.reshape(-1, out.shape[-1])
    return gemm_desc([a2], [(0, 0, 0, 0, ceil64(k) // 64)], w, o2, w.shape[0], **kw)


def linear2_desc(a0: torch.Tensor, a1: torch.Tensor, w: torch.Tensor, out: torch.Tensor, **kw) -> L.GemmDesc:
    """1x1 conv over the channel concat [a0 | a1] without materialising it (two K segments)."""
    k0, k1 = a0.shape[-1], a1.shape[-1]
    assert k0 % 64 == 0, "first concat source must be a multiple of 64 channels"
    f = lambda t: t if t.dim() == 2 else t.reshape(-1, t.shape[-1])
    return gemm_desc([f(a0), f(a1)], [(0, 0, 0, 0, k0 // 64), (1, 0, 0, 0, ceil64(k1) // 64)], w, f(out),
                     w.shape[0], **kw)


def conv3x3_desc(x: torch.Tensor, wp: torch.Tensor, out: torch.Tensor, **kw) -> L.GemmDesc:
    """3x3 stride-1 pad-1 conv. x [B,H,W,Cin], wp packed [Cout, 9*ceil64(Cin)], out [B,H,W,Cout]."""
    cin = x.shape[-1]
    nch = ceil64(cin) // 64
    segs = [(0, 0, kx - 1, ky - 1, nch) for ky in range(3) for kx in range(3)]
    return gemm_desc([x], segs, wp, out, wp.shape[0], **kw)


def conv3x3_halo_desc(xs: Sequence[torch.Tensor], wp: torch.Tensor, out: torch.Tensor, gn: Optional[tuple] = None,
                      extra_srcs: Sequence[torch.Tensor] = (), extra_segs: Sequence[tuple] = (), **kw) -> L.GemmDesc:
    """3x3 stride-1 pad-1 conv over the channel concat of ``xs`` (one or two [B,H,W,C_i] tensors) in the halo-tile
    form (ldx_gemm_desc_t.halo_nsrc): every 64-channel chunk is fetched once as a 10 x 18 halo tile and read by all
    nine taps.  wp: pack_conv3x3 of the conv weight over the concatenated channels, optionally followed by the columns
    of extra 1x1 K segments (``extra_srcs`` / ``extra_segs`` = (index into extra_srcs, c0, dx, dy, nchunks)).
    gn = (coef [B, Ctot, 2] fp32 from group_norm_coef_op, silu): GroupNorm (+ SiLU) applied to the landed tiles."""
    assert 1 <= len(xs) <= 2
    srcs = list(xs) + list(extra_srcs)
    segs = [(len(xs) + si, c0, dx, dy, nch) for (si, c0, dx, dy, nch) in extra_segs]
    d = gemm_desc(srcs, segs, wp, out, wp.shape[0], **kw)
    d.halo_nsrc = len(xs)
    if gn is not None:
        coef, silu = gn
        ctot = sum(t.shape[-1] for t in xs)
        assert coef.dtype == torch.float32 and coef.is_contiguous() and tuple(coef.shape) == (xs[0].shape[0], ctot, 2)
        d.gn_coef, d.gn_ctot, d.gn_silu = coef.data_ptr(), ctot, int(bool(silu))
    return d


def conv1x1_gn_desc(x: torch.Tensor, w: torch.Tensor, out: torch.Tensor, gn: tuple, **kw) -> L.GemmDesc:
    """1x1 conv (a Linear over the channels) of x [B,H,W,C] with GroupNorm (+ SiLU) applied to the operand tiles in shared
    memory (SpatialTransformer.norm + proj_in, LD.py:4246-4251); out [B,H,W,N].  gn = (coef [B, C, 2], silu)."""
    assert x.dim() == 4 and out.dim() == 4 and x.shape[:3] == out.shape[:3]
    c = x.shape[-1]
    d = gemm_desc([x], [(0, 0, 0, 0, ceil64(c) // 64)], w, out, w.shape[0], **kw)
    coef, silu = gn
    assert coef.dtype == torch.float32 and coef.is_contiguous() and tuple(coef.shape) == (x.shape[0], c, 2)
    d.gn_coef, d.gn_ctot, d.gn_silu = coef.data_ptr(), c, int(bool(silu))
    return d


def halo_ok(x: torch.Tensor, out: torch.Tensor) -> bool:
    """Shapes the halo-tile conv serves well: 8 x 16-pixel tiles that the map fills (W % 8 == 0, H % 16 == 0)."""
    _, h, w, _ = x.shape
    return w % 8 == 0 and h % 16 == 0 and out.dtype == torch.float16 and out.stride(2) % 8 == 0


HALO_GENERIC = os.environ.get("LDX_HALO_GENERIC", "1") == "1"


def conv3x3_auto_desc(x: torch.Tensor, wp: torch.Tensor, out: torch.Tensor, extra_srcs: Sequence[torch.Tensor] = (),
                      extra_segs: Sequence[tuple] = (), **kw) -> L.GemmDesc:
    """3x3 stride-1 pad-1 conv (+ optional extra 1x1 K segments whose weight columns follow the nine taps') through the
    halo-tile kernel when the map tiles into 8 x 16 pixels, else through the implicit GEMM."""
    if HALO_GENERIC and halo_ok(x, out) and x.shape[-1] % 8 == 0:
        return conv3x3_halo_desc([x], wp, out, extra_srcs=extra_srcs, extra_segs=extra_segs, **kw)
    nch = ceil64(x.shape[-1]) // 64
    segs = [(0, 0, kx - 1, ky - 1, nch) for ky in range(3) for kx in range(3)]
    segs += [(1 + si, c0, dx, dy, n) for (si, c0, dx, dy, n) in extra_segs]
    return gemm_desc([x] + list(extra_srcs), segs, wp, out, wp.shape[0], **kw)


def conv3x3_s2_desc(x: torch.Tensor, wp: torch.Tensor, out: torch.Tensor, pad_end: bool = False, **kw) -> L.GemmDesc:
    """3x3 stride-2 conv via four parity-strided views of x: pad 1 on every side (Downsample1, LD.py:5173-5186) or,
    with ``pad_end``, the VAE encoder's F.pad(x, (0,1,0,1)) + padding-0 conv (Downsample, LD.py:3514-3528), where
    tap k reads input row 2y + k: parity k & 1, shifted by k >> 1."""
    b, h, w, cin = x.shape
    assert h % 2 == 0 and w % 2 == 0
    nch = ceil64(cin) // 64
    views = [x[:, py::2, px::2, :] for py in range(2) for px in range(2)]
    taps = ((0, 0), (1, 0), (0, 1)) if pad_end else ((1, -1), (0, 0), (1, 0))
    segs = []
    for ky in range(3):
        py, dy = taps[ky]
        for kx in range(3):
            px, dx = taps[kx]
            segs.append((py * 2 + px, 0, dx, dy, nch))
    return gemm_desc(views, segs, wp, out, wp.shape[0], **kw)


def upconv3x3_descs(x: torch.Tensor, wp4: torch.Tensor, out: torch.Tensor, **kw) -> List[L.GemmDesc]:
    """nearest-2x upsample + 3x3 conv as four 2x2 phase convs on the low-res input.
    x [B,H,W,Cin], wp4 packed [4, Cout, 4*ceil64(Cin)], out [B,2H,2W,Cout]."""
    cin = x.shape[-1]
    nch = ceil64(cin) // 64
    descs = []
    for py in range(2):
        for px in range(2):
            segs = []
            for a in range(2):
                dy = a - 1 if py == 0 else a
                for bb in range(2):
                    dx = bb - 1 if px == 0 else bb
                    segs.append((0, 0, dx, dy, nch))
            descs.append(gemm_desc([x], segs, wp4[py * 2 + px], out[:, py::2, px::2, :], wp4.shape[1], **kw))
    return descs


def run_gemm(d: L.GemmDesc):
    L.check(L.lib().ldx_gemm(C.byref(d), _stream()), "ldx_gemm")


# ----------------------------------------------------------------------------- weight packing
def pack_conv3x3(w_oihw: torch.Tensor) -> torch.Tensor:
    cout, cin = w_oihw.shape[:2]
    w16 = _dev16(w_oihw)
    out = torch.empty(cout, 9 * ceil64(cin), dtype=torch.float16, device="cuda")
    L.check(L.lib().ldx_pack_conv3x3(w16.data_ptr(), out.data_ptr(), cout, cin, _stream()), "ldx_pack_conv3x3")
    return out


def pack_upconv3x3(w_oihw: torch.Tensor) -> torch.Tensor:
    cout, cin = w_oihw.shape[:2]
    w16 = _dev16(w_oihw)
    out = torch.empty(4, cout, 4 * ceil64(cin), dtype=torch.float16, device="cuda")
    L.check(L.lib().ldx_pack_upconv3x3(w16.data_ptr(), out.data_ptr(), cout, cin, _stream()), "ldx_pack_upconv3x3")
    return out


def pack_conv3x3_small(w_oihw: torch.Tensor) -> torch.Tensor:
    """[Cout <= 5, Cin, 3, 3] -> the [ceil16(9 * Cout), ceil64(Cin)] per-tap weight matrix of ldx_conv3x3_small."""
    cout, cin = w_oihw.shape[:2]
    n = L.lib().ldx_conv3x3_small_packed_elems(cout, cin)
    if n < 0:
        raise L.LdxError(f"conv3x3_small: Cout {cout} unsupported")
    w16 = _dev16(w_oihw)
    out = torch.empty(n // ceil64(cin), ceil64(cin), dtype=torch.float16, device="cuda")
    L.check(L.lib().ldx_pack_conv3x3_small(w16.data_ptr(), out.data_ptr(), cout, cin, _stream()), "ldx_pack_conv3x3_small")
    return out


def conv3x3_small_op(x: torch.Tensor, wp: torch.Tensor, bias: Optional[torch.Tensor], cout: int,
                     out16: Optional[torch.Tensor] = None, out_img: Optional[torch.Tensor] = None,
                     post: int = 0) -> L.GenericOp:
    """x [B,H,W,Cin] fp16 (pixel-contiguous, channel pitch free) -> out16 [B,H,W,>=cout] fp16 and / or out_img
    [B,H,W,cout] fp32 (post = 1: clamp((v + 1) / 2, 0, 1), the VAE's process_output)."""
    _check_act(x)
    b, h, w, cin = x.shape
    assert x.stride(1) == x.stride(2) * w and x.stride(0) == x.stride(1) * h, "input must be pixel-contiguous"
    if out16 is not None:
        assert out16.dtype == torch.float16 and out16.stride(-1) == 1 and out16.shape[:3] == x.shape[:3]
        assert out16.stride(1) == out16.stride(2) * w and out16.stride(0) == out16.stride(1) * h
    if out_img is not None:
        assert out_img.dtype == torch.float32 and out_img.is_contiguous() and tuple(out_img.shape) == (b, h, w, cout)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == cout
    return generic_op(L.OP_CONV_SMALL, p=[x.data_ptr(), wp.data_ptr(), _ptr(bias), _ptr(out16), _ptr(out_img)],
                      i=[b, h, w, cin, cout, post], l=[x.stride(2), 0 if out16 is None else out16.stride(2)])


def pack_linear(w: torch.Tensor) -> torch.Tensor:
    """[N, K] (or 1x1 conv [N, K, 1, 1]) -> fp16 [N, ceil64(K)] (zero padded)."""
    w2 = w.detach().reshape(w.shape[0], -1).to(torch.float16)
    n, k = w2.shape
    if k % 64 == 0:
        return w2.contiguous().to("cuda")
    out = torch.zeros(n, ceil64(k), dtype=torch.float16, device=w2.device)
    out[:, :k] = w2
    return out.to("cuda")


def pack_geglu(w: torch.Tensor, bias: torch.Tensor, half_bn: int = 128):
    w16 = _dev16(w)
    b32 = bias.detach().to(torch.float32).contiguous().to("cuda")
    two_f, k = w16.shape
    assert k % 64 == 0
    wo, bo = torch.empty_like(w16), torch.empty_like(b32)
    L.check(L.lib().ldx_pack_geglu(w16.data_ptr(), b32.data_ptr(), wo.data_ptr(), bo.data_ptr(), two_f, k, half_bn,
                                   _stream()), "ldx_pack_geglu")
    return wo, bo


# ----------------------------------------------------------------------------- generic ops
def generic_op(kind: int, p=(), i=(), l=(), f=()) -> L.GenericOp:
    op = L.GenericOp()
    op.kind = kind
    for k, v in enumerate(p):
        op.p[k] = v
    for k, v in enumerate(i):
        op.i[k] = v
    for k, v in enumerate(l):
        op.l[k] = v
    for k, v in enumerate(f):
        op.f[k] = v
    return op


def group_norm_op(x0, x1, gamma, beta, y, stats, eps: float, silu: bool, colstats=None) -> L.GenericOp:
    """x0 (and optional x1, channel-concatenated) [B,H,W,C*] -> y [B,H,W,C0+C1].
    colstats = ((buf0, slots0), (buf1, slots1) | None): column partials written by the GEMMs that produced x0 / x1
    (attach_colstats); the statistics pass over the tensor is then replaced by a reduction over them."""
    b, h, w, c0 = x0.shape
    assert x0.stride(1) == x0.stride(2) * w and x0.stride(0) == x0.stride(1) * h
    c1 = 0 if x1 is None else x1.shape[-1]
    assert stats.numel() >= b * L.GN_STATS_FLOATS_PER_BATCH and stats.dtype == torch.float32
    cs0 = cs1 = None
    s0 = s1 = 0
    if colstats is not None and colstats[0] is not None and (x1 is None or colstats[1] is not None):
        (cs0, s0) = colstats[0]
        assert tuple(cs0.shape) == (b, s0, c0, 2) and cs0.dtype == torch.float32 and cs0.is_contiguous()
        if x1 is not None:
            (cs1, s1) = colstats[1]
            assert tuple(cs1.shape) == (b, s1, c1, 2) and cs1.dtype == torch.float32 and cs1.is_contiguous()
    return generic_op(L.OP_GROUP_NORM, p=[x0.data_ptr(), _ptr(x1), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(),
                                          stats.data_ptr(), _ptr(cs0), _ptr(cs1)],
                      i=[c0, c1, b, h * w, int(silu), s0, s1],
                      l=[x0.stride(2), 0 if x1 is None else x1.stride(2), y.stride(2)], f=[eps])


def group_norm_coef_op(x0, x1, gamma, beta, coef, stats, tickets, eps: float, colstats=None) -> L.GenericOp:
    """GroupNorm statistics of x0 (and optional x1, channel-concatenated) as per-(sample, channel) coefficients
    coef [B, C0+C1, 2] fp32 = (rstd * gamma, beta - mean * rstd * gamma) for conv3x3_halo_desc(gn=...): the normalisation
    itself happens in the consuming conv's operand path.  tickets: B zero-initialised int32 (left zeroed)."""
    b, h, w, c0 = x0.shape
    assert x0.stride(1) == x0.stride(2) * w and x0.stride(0) == x0.stride(1) * h
    c1 = 0 if x1 is None else x1.shape[-1]
    assert stats.numel() >= b * L.GN_STATS_FLOATS_PER_BATCH and stats.dtype == torch.float32
    assert coef.dtype == torch.float32 and coef.is_contiguous() and tuple(coef.shape) == (b, c0 + c1, 2)
    assert tickets.dtype == torch.int32 and tickets.numel() >= b
    cs0 = cs1 = None
    s0 = s1 = 0
    if colstats is not None and colstats[0] is not None and (x1 is None or colstats[1] is not None):
        (cs0, s0) = colstats[0]
        assert tuple(cs0.shape) == (b, s0, c0, 2) and cs0.dtype == torch.float32 and cs0.is_contiguous()
        if x1 is not None:
            (cs1, s1) = colstats[1]
            assert tuple(cs1.shape) == (b, s1, c1, 2) and cs1.dtype == torch.float32 and cs1.is_contiguous()
    return generic_op(L.OP_GROUP_NORM_COEF, p=[x0.data_ptr(), _ptr(x1), gamma.data_ptr(), beta.data_ptr(), coef.data_ptr(),
                                               stats.data_ptr(), _ptr(cs0), _ptr(cs1), tickets.data_ptr()],
                      i=[c0, c1, b, h * w, 0, s0, s1],
                      l=[x0.stride(2), 0 if x1 is None else x1.stride(2)], f=[eps])


def gn_fold_weights_op(w: torch.Tensor, bias: Optional[torch.Tensor], beta: torch.Tensor, coef: torch.Tensor,
                       w_out: torch.Tensor, bias_out: torch.Tensor) -> L.GenericOp:
    """GroupNorm folded into the linear / 1x1 conv that consumes it (ldx_gn_fold_weights): w [N, ldw] fp16 (pack_linear),
    coef [B, C, 2] from group_norm_coef_op, w_out [B, N, ldw] fp16, bias_out [B, N] fp32."""
    B, C = coef.shape[0], coef.shape[1]
    N = w.shape[0]
    assert w.dtype == torch.float16 and w.stride(1) == 1 and w_out.is_contiguous() and tuple(w_out.shape) == (B, N, w.stride(0))
    assert bias_out.dtype == torch.float32 and bias_out.is_contiguous() and tuple(bias_out.shape) == (B, N)
    assert coef.dtype == torch.float32 and coef.is_contiguous() and beta.dtype == torch.float32 and beta.numel() == C
    return generic_op(L.OP_GN_FOLD_WEIGHTS, p=[_ptr(w), _ptr(bias), _ptr(beta), _ptr(coef), _ptr(w_out), _ptr(bias_out)],
                      l=[w.stride(0)], i=[B, N, C])


def conv1x1_batched_desc(x: torch.Tensor, w_b: torch.Tensor, bias_b: torch.Tensor, out: torch.Tensor, **kw) -> L.GemmDesc:
    """1x1 conv with per-sample weights and bias (ldx_gemm_desc_t.w_batch_rows): x [B, H, W, C], w_b [B, N, ldw] fp16 and
    bias_b [B, N] fp32 from gn_fold_weights_op, out [B, H, W, N]."""
    B, N, ldw = w_b.shape
    assert x.dim() == 4 and x.shape[0] == B and bias_b.dtype == torch.float32 and tuple(bias_b.shape) == (B, N)
    d = gemm_desc([x], [(0, 0, 0, 0, ceil64(x.shape[-1]) // 64)], w_b.view(B * N, ldw), out, N, bias=bias_b.view(-1), **kw)
    d.w_batch_rows, d.bias_batch_stride = N, N
    return d


def ff_fused_op(x2, w1p, bias1, gsum1, w2, bias2, rowstats, out2, eps=1e-5) -> L.GenericOp:
    """out = x + W2 . GEGLU(LayerNorm(x) . W1^T + b1) + b2 in one launch (C = 320).  w1p / bias1 / gsum1: the LayerNorm-folded
    GEGLU projection packed with pack_geglu(half_bn=64); rowstats [M, parts, 2]: row statistics of x from its producer."""
    m, c = x2.shape
    f = w2.shape[1]
    assert x2.dtype == torch.float16 and out2.dtype == torch.float16 and tuple(out2.shape) == (m, c)
    assert w1p.shape[0] == 2 * f and w1p.stride(1) == 1 and w2.shape[0] == c and w2.stride(1) == 1
    assert rowstats.dtype == torch.float32 and rowstats.is_contiguous() and rowstats.shape[0] == m and rowstats.shape[2] == 2
    return generic_op(L.OP_FF_FUSED, p=[x2.data_ptr(), w1p.data_ptr(), bias1.data_ptr(), gsum1.data_ptr(), w2.data_ptr(),
                                        bias2.data_ptr(), rowstats.data_ptr(), out2.data_ptr()],
                      l=[x2.stride(0), w1p.stride(0), w2.stride(0), out2.stride(0)], i=[m, c, f, rowstats.shape[1]], f=[eps])


def layer_norm_op(x2, gamma, beta, y2, eps=1e-5) -> L.GenericOp:
    rows, c = x2.shape
    return generic_op(L.OP_LAYER_NORM, p=[x2.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y2.data_ptr()],
                      i=[rows, c], l=[x2.stride(0), y2.stride(0)], f=[eps])


def attention_op(q, k, v, o, heads: int, causal: bool = False) -> L.GenericOp:
    """q [B,Nq,heads*d], k/v [B,Nk,heads*d] (row pitch free, batch stride = N * pitch), o like q."""
    b, nq, c = q.shape
    nk = k.shape[1]
    for t, n in ((q, nq), (k, nk), (v, nk), (o, nq)):
        _check_act(t)
        assert t.stride(0) == t.stride(1) * n, "batch stride must equal N * row pitch"
    return generic_op(L.OP_ATTENTION, p=[q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr()],
                      l=[q.stride(1), k.stride(1), v.stride(1), o.stride(1)], i=[b, nq, nk, heads, c // heads, 1 if causal else 0])


def run_op(op: L.GenericOp):
    """Execute one generic op immediately (through a one-op program)."""
    prog = Program()
    prog.add(op)
    prog.run()


class Program:
    """Recorded list of kernel launches (ldx_program_*)."""

    def __init__(self):
        L.require_device()
        self._h = L.lib().ldx_program_create()
        self._keep = []          # tensors referenced by raw pointers
        self.meta = []           # per op: (kind, executed flops, tag) for profiling

    def add(self, op, keep=()):
        if isinstance(op, L.GemmDesc):
            L.check(L.lib().ldx_program_add_gemm(self._h, C.byref(op)), "ldx_program_add_gemm")
            k = 64 * sum(op.seg[i].nchunks for i in range(op.nseg))
            k += 9 * sum(ceil64(op.src[i].C) for i in range(op.halo_nsrc))
            kind = "conv_halo" if (op.halo_nsrc > 0 or op.gn_coef) else "gemm"      # which kernel runs it
            self.meta.append((kind, 2.0 * op.W * op.H * op.B * op.N * k, (op.W * op.H * op.B, op.N, k)))
        elif isinstance(op, (list, tuple)):
            for o in op:
                self.add(o)
        else:
            L.check(L.lib().ldx_program_add_op(self._h, C.byref(op)), "ldx_program_add_op")
            flops, tag = 0.0, None
            if op.kind == L.OP_ATTENTION:
                b, nq, nk, heads, d = (op.i[j] for j in range(5))
                flops, tag = 4.0 * b * nq * nk * heads * d, (b, nq, nk, heads, d)
            elif op.kind == L.OP_FF_FUSED:
                m, c, f = op.i[0], op.i[1], op.i[2]
                flops, tag = 2.0 * m * c * 3 * f, (m, c, f)
            elif op.kind == L.OP_CONV_SMALL:
                b, h, w, cin, cout = (op.i[j] for j in range(5))
                flops, tag = 2.0 * b * h * w * 9 * cin * cout, (b * h * w, cout, 9 * cin)
            elif op.kind in (L.OP_GROUP_NORM, L.OP_GROUP_NORM_COEF):
                tag = (op.i[2], op.i[3], op.i[0] + op.i[1])         # B, HW, C
            elif op.kind == L.OP_LAYER_NORM:
                tag = (op.i[0], op.i[1])                            # rows, C
            self.meta.append(({L.OP_ATTENTION: "attention", L.OP_GROUP_NORM: "group_norm", L.OP_FF_FUSED: "ff_fused",
                               L.OP_GROUP_NORM_COEF: "group_norm", L.OP_GN_FOLD_WEIGHTS: "group_norm",
                               L.OP_LAYER_NORM: "layer_norm", L.OP_CONV_SMALL: "conv_small"}.get(op.kind, "other"),
                              flops, tag))
        self._keep.extend(keep)

    def keep(self, *tensors):
        self._keep.extend(tensors)

    @property
    def num_launches(self) -> int:
        return L.lib().ldx_program_num_launches(self._h)

    def run(self):
        L.check(L.lib().ldx_program_run(self._h, _stream()), "ldx_program_run")

    def run_timed(self):
        """Blocking replay with per-op device times; returns [(kind, flops, tag, ms)]."""
        n = L.lib().ldx_program_num_ops(self._h)
        ms = (C.c_float * n)()
        L.check(L.lib().ldx_program_run_timed(self._h, _stream(), ms, n), "ldx_program_run_timed")
        return [(m[0], m[1], m[2], ms[i]) for i, m in enumerate(self.meta)]

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                L.lib().ldx_program_destroy(h)
            except Exception:
                pass
