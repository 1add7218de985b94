"""Kernel-level parity: every libldx.so kernel against the torch op it replaces, computed in fp32
from the same fp16 inputs (the oracle for a floating-point kernel).  All calls go through the C ABI.
Tolerances: fp16 output rounding (rel 2^-11) plus fp32 accumulation-order noise -> 2e-3 rel-L2 bound
per kernel is loose; observed values are ~3e-4."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from lightdiffusion_b200 import _lib as L
    from lightdiffusion_b200 import kernels as K
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

DEV = "cuda"


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def h16(t):
    return t.to(torch.float16)


# ------------------------------------------------------------------------------- GEMM / linear
@pytest.mark.parametrize("M,N,Kd", [(128, 128, 64), (256, 320, 320), (4096, 320, 320), (77 * 2, 640, 768),
                                    (16, 1280, 320), (1000, 2560, 1280), (8192, 1280, 5120), (130, 72, 128)])
def test_linear_plain(M, N, Kd):
    a, w = h16(rnd(M, Kd, seed=1)), h16(rnd(N, Kd, seed=2, scale=Kd ** -0.5))
    out = torch.empty(M, N, dtype=torch.float16, device=DEV)
    K.run_gemm(K.linear_desc(a, w, out))
    ref = a.float() @ w.float().t()
    assert rel(out, ref) < 1e-3, (M, N, Kd, rel(out, ref))


@pytest.mark.parametrize("bn", [32, 64, 96, 128, 160, 192, 224, 256])
def test_linear_all_tile_widths(bn):
    M, N, Kd = 512, 640, 448
    a, w = h16(rnd(M, Kd, seed=3)), h16(rnd(N, Kd, seed=4, scale=Kd ** -0.5))
    out = torch.empty(M, N, dtype=torch.float16, device=DEV)
    K.run_gemm(K.linear_desc(a, w, out, bn=bn))
    assert rel(out, a.float() @ w.float().t()) < 1e-3


def test_linear_epilogues():
    M, N, Kd = 1024, 320, 320
    a, w = h16(rnd(M, Kd, seed=5)), h16(rnd(N, Kd, seed=6, scale=Kd ** -0.5))
    bias = rnd(N, seed=7)
    res = h16(rnd(M, N, seed=8))
    out = torch.empty(M, N, dtype=torch.float16, device=DEV)
    K.run_gemm(K.linear_desc(a, w, out, bias=bias, residual=res))
    ref = a.float() @ w.float().t() + bias + res.float()
    assert rel(out, ref) < 1e-3
    # SiLU + fp32 output
    out32 = torch.empty(M, N, dtype=torch.float32, device=DEV)
    K.run_gemm(K.linear_desc(a, w, out32, bias=bias, flags=L.EPI_SILU | L.EPI_OUT_F32))
    assert rel(out32, F.silu(a.float() @ w.float().t() + bias)) < 1e-5
    # strided output / input views (channel slices of wider buffers)
    wide_in = h16(rnd(M, Kd + 64, seed=9))
    wide_out = torch.zeros(M, N + 32, dtype=torch.float16, device=DEV)
    K.run_gemm(K.linear_desc(wide_in[:, 64:], w, wide_out[:, 32:]))
    assert rel(wide_out[:, 32:], wide_in[:, 64:].float() @ w.float().t()) < 1e-3
    assert float(wide_out[:, :32].abs().max()) == 0.0


def test_linear_scalar_tail_n4():
    M, N, Kd = 4096, 4, 320      # UNet out conv has N=4 -> scalar store path
    a, w = h16(rnd(M, Kd, seed=10)), h16(rnd(N, Kd, seed=11, scale=Kd ** -0.5))
    out = torch.zeros(M, 8, dtype=torch.float16, device=DEV)
    K.run_gemm(K.linear_desc(a, w, out[:, :4]))
    assert rel(out[:, :4], a.float() @ w.float().t()) < 1e-3
    assert float(out[:, 4:].abs().max()) == 0.0


@pytest.mark.parametrize("M,Kd,Fd,half_bn", [(512, 320, 1280, 128), (300, 640, 2560, 128), (256, 1280, 5120, 64)])
def test_geglu(M, Kd, Fd, half_bn):
    a = h16(rnd(M, Kd, seed=12))
    w = h16(rnd(2 * Fd, Kd, seed=13, scale=Kd ** -0.5))
    bias = rnd(2 * Fd, seed=14, scale=0.3)
    wp, bp = K.pack_geglu(w, bias, half_bn)
    out = torch.empty(M, Fd, dtype=torch.float16, device=DEV)
    K.run_gemm(K.linear_desc(a, wp, out, bias=bp, flags=L.EPI_GEGLU, bn=2 * half_bn))
    y = a.float() @ w.float().t() + bias
    ref = y[:, :Fd] * F.gelu(y[:, Fd:])
    assert rel(out, ref) < 1e-3


@pytest.mark.parametrize("M,C,N,geglu", [(1024, 320, 960, False), (4096 + 77, 640, 640, False), (512, 1280, 1280, False),
                                         (2048, 320, 2560, True), (300, 640, 5120, True)])
def test_layer_norm_folded_into_gemm(M, C, N, geglu):
    """x = a @ w0^T + b0 + r (producer, leaves per-row statistics) ; out = Linear(LayerNorm(x)) with the
    normalisation folded into the consumer GEMM (LD.py:4117-4162: x += attn(norm(x)), ff(norm3(x)))."""
    K0 = 320
    a, w0, b0 = h16(rnd(M, K0, seed=30)), h16(rnd(C, K0, seed=31, scale=K0 ** -0.5)), rnd(C, seed=32)
    r = h16(rnd(M, C, seed=33) + 0.7)                 # non-zero row mean
    x = torch.empty(M, C, dtype=torch.float16, device=DEV)
    d0 = K.linear_desc(a, w0, x, bias=b0, residual=r)
    stats = K.attach_rowstats(d0, M)
    K.run_gemm(d0)
    xs = x.float()
    assert rel(stats[:, :, 0].sum(1), xs.sum(1)) < 2e-3
    assert rel(stats[:, :, 1].sum(1), (xs * xs).sum(1)) < 2e-3

    gamma, beta = 1.0 + 0.2 * rnd(C, seed=34), 0.3 * rnd(C, seed=35)
    w, bias = h16(rnd(N, C, seed=36, scale=C ** -0.5)), rnd(N, seed=37, scale=0.3)
    wf, gsum, bf = K.fold_layer_norm(w, bias, gamma, beta)
    y = F.layer_norm(xs, (C,), gamma, beta, 1e-5) @ w.float().t() + bias
    if geglu:
        wp, bp = K.pack_geglu(wf, bf, 128)
        _, gp = K.pack_geglu(wf, gsum, 128)
        out = torch.empty(M, N // 2, dtype=torch.float16, device=DEV)
        K.run_gemm(K.linear_desc(x, wp, out, bias=bp, flags=L.EPI_GEGLU, bn=256, rownorm=(stats, gp, 1e-5, C)))
        ref = y[:, :N // 2] * F.gelu(y[:, N // 2:])
    else:
        out = torch.empty(M, N, dtype=torch.float16, device=DEV)
        K.run_gemm(K.linear_desc(x, wf, out, bias=bf, rownorm=(stats, gsum, 1e-5, C)))
        ref = y
    assert rel(out, ref) < 1.5e-3, rel(out, ref)


@pytest.mark.parametrize("Cin,res,silu", [(1280, True, False), (2560, False, False), (640, False, True)])
def test_conv3x3_split_k(Cin, res, silu):
    """8x8 maps (M = 1024): the library splits K over CTA-pair groups and reduces fp32 partials in a second
    launch that also applies bias / timestep row add / residual / SiLU."""
    B, H, W, Cout = 16, 8, 8, 1280
    x = h16(rnd(B, H, W, Cin, seed=40))
    w = h16(rnd(Cout, Cin, 3, 3, seed=41, scale=(9 * Cin) ** -0.5))
    bias, temb = rnd(Cout, seed=42), rnd(B, Cout, seed=43)
    r = h16(rnd(B, H, W, Cout, seed=44)) if res else None
    out = torch.empty(B, H, W, Cout, dtype=torch.float16, device=DEV)
    K.run_gemm(K.conv3x3_desc(x, K.pack_conv3x3(w), out, bias=bias, rowadd=temb, residual=r,
                              flags=L.EPI_SILU if silu else 0))
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1) + temb[:, :, None, None]
    if res:
        ref = ref + r.float().permute(0, 3, 1, 2)
    if silu:
        ref = F.silu(ref)
    assert rel(out.permute(0, 3, 1, 2), ref) < 1e-3


def test_rowadd_timestep_broadcast():
    B, H, W, Cin, Cout = 3, 16, 16, 64, 96
    x = h16(rnd(B, H, W, Cin, seed=15))
    w = h16(rnd(Cout, Cin, 3, 3, seed=16, scale=(9 * Cin) ** -0.5))
    bias, temb = rnd(Cout, seed=17), rnd(B, Cout + 5, seed=18)
    out = torch.empty(B, H, W, Cout, dtype=torch.float16, device=DEV)
    K.run_gemm(K.conv3x3_desc(x, K.pack_conv3x3(w), out, bias=bias, rowadd=temb[:, 5:]))
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1) + temb[:, 5:, None, None]
    assert rel(out.permute(0, 3, 1, 2), ref) < 1e-3


# ------------------------------------------------------------------------------- convolutions
@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 64, 64, 320, 320), (2, 32, 32, 640, 640), (4, 16, 16, 1280, 1280),
                                            (3, 8, 8, 1280, 1280), (1, 8, 8, 64, 64), (2, 64, 64, 8, 320),
                                            (1, 24, 40, 128, 64), (2, 128, 128, 128, 128), (2, 64, 64, 320, 4)])
def test_conv3x3(B, H, W, Cin, Cout):
    x = h16(rnd(B, H, W, Cin, seed=20))
    w = h16(rnd(Cout, Cin, 3, 3, seed=21, scale=(9 * Cin) ** -0.5))
    bias = rnd(Cout, seed=22)
    res = h16(rnd(B, H, W, Cout, seed=23))
    out = torch.empty(B, H, W, Cout, dtype=torch.float16, device=DEV)
    K.run_gemm(K.conv3x3_desc(x, K.pack_conv3x3(w), out, bias=bias, residual=res))
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1) + res.float().permute(0, 3, 1, 2)
    assert rel(out.permute(0, 3, 1, 2), ref) < 1e-3, rel(out.permute(0, 3, 1, 2), ref)


@pytest.mark.parametrize("B,H,W,C", [(2, 64, 64, 320), (2, 32, 32, 640), (3, 16, 16, 1280)])
def test_conv3x3_stride2(B, H, W, C):
    x = h16(rnd(B, H, W, C, seed=24))
    w = h16(rnd(C, C, 3, 3, seed=25, scale=(9 * C) ** -0.5))
    bias = rnd(C, seed=26)
    out = torch.empty(B, H // 2, W // 2, C, dtype=torch.float16, device=DEV)
    K.run_gemm(K.conv3x3_s2_desc(x, K.pack_conv3x3(w), out, bias=bias))
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, stride=2, padding=1)
    assert rel(out.permute(0, 3, 1, 2), ref) < 1e-3


@pytest.mark.parametrize("B,H,W,C", [(2, 64, 48, 128), (1, 32, 32, 256), (2, 16, 24, 512)])
def test_conv3x3_stride2_pad_end(B, H, W, C):
    """The VAE encoder's Downsample (LD.py:3514-3528): F.pad(x, (0,1,0,1)) then a stride-2, padding-0 conv."""
    x = h16(rnd(B, H, W, C, seed=124))
    w = h16(rnd(C, C, 3, 3, seed=125, scale=(9 * C) ** -0.5))
    bias = rnd(C, seed=126)
    out = torch.empty(B, H // 2, W // 2, C, dtype=torch.float16, device=DEV)
    K.run_gemm(K.conv3x3_s2_desc(x, K.pack_conv3x3(w), out, pad_end=True, bias=bias))
    ref = F.conv2d(F.pad(x.float().permute(0, 3, 1, 2), (0, 1, 0, 1)), w.float(), bias, stride=2)
    assert rel(out.permute(0, 3, 1, 2), ref) < 1e-3


@pytest.mark.parametrize("B,H,W,C", [(2, 8, 8, 1280), (2, 16, 16, 1280), (2, 32, 32, 640), (1, 64, 64, 128)])
def test_upsample_conv_folded(B, H, W, C):
    x = h16(rnd(B, H, W, C, seed=27))
    w = h16(rnd(C, C, 3, 3, seed=28, scale=(9 * C) ** -0.5))
    bias = rnd(C, seed=29)
    out = torch.empty(B, 2 * H, 2 * W, C, dtype=torch.float16, device=DEV)
    for d in K.upconv3x3_descs(x, K.pack_upconv3x3(w), out, bias=bias):
        K.run_gemm(d)
    up = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    ref = F.conv2d(up, w.float(), bias, padding=1)
    # folding sums up to 4 fp16 weights before rounding -> slightly different rounding than the reference
    assert rel(out.permute(0, 3, 1, 2), ref) < 1.5e-3
    # explicit upsample kernel
    y = torch.empty(B, 2 * H, 2 * W, C, dtype=torch.float16, device=DEV)
    L.check(L.lib().ldx_upsample_nearest2x(x.data_ptr(), y.data_ptr(), B, H, W, C, 0))
    torch.cuda.synchronize()
    assert torch.equal(y.permute(0, 3, 1, 2).float(), up)


def test_concat_1x1_skip():
    B, H, W, C0, C1, Cout = 2, 16, 16, 1280, 640, 1280
    a0, a1 = h16(rnd(B, H, W, C0, seed=30)), h16(rnd(B, H, W, C1, seed=31))
    w = h16(rnd(Cout, C0 + C1, seed=32, scale=(C0 + C1) ** -0.5))
    bias = rnd(Cout, seed=33)
    out = torch.empty(B, H, W, Cout, dtype=torch.float16, device=DEV)
    K.run_gemm(K.linear2_desc(a0, a1, w, out, bias=bias))
    ref = torch.cat([a0, a1], -1).float() @ w.float().t() + bias
    assert rel(out, ref) < 1e-3


# ------------------------------------------------------------------------------- norms
@pytest.mark.parametrize("B,HW,C0,C1,silu,eps", [(2, 4096, 320, 0, True, 1e-5), (2, 1024, 1280, 640, True, 1e-5),
                                                 (3, 64, 1280, 1280, True, 1e-5), (2, 4096, 640, 320, True, 1e-5),
                                                 (2, 256, 1280, 0, False, 1e-6), (1, 16384, 128, 0, True, 1e-6),
                                                 (2, 4096, 512, 0, False, 1e-6)])
def test_group_norm(B, HW, C0, C1, silu, eps):
    h = int(math.isqrt(HW))
    x0 = h16(rnd(B, h, HW // h, C0, seed=40) * 2 + 0.7)
    x1 = h16(rnd(B, h, HW // h, C1, seed=41) * 0.5 - 0.3) if C1 else None
    C = C0 + C1
    gamma, beta = rnd(C, seed=42) * 0.2 + 1, rnd(C, seed=43) * 0.2
    y = torch.empty(B, h, HW // h, C, dtype=torch.float16, device=DEV)
    stats = torch.empty(B * L.GN_STATS_FLOATS_PER_BATCH, dtype=torch.float32, device=DEV)
    K.run_op(K.group_norm_op(x0, x1, gamma, beta, y, stats, eps, silu))
    xc = (x0 if x1 is None else torch.cat([x0, x1], -1)).float().permute(0, 3, 1, 2)
    ref = F.group_norm(xc, 32, gamma, beta, eps)
    ref = F.silu(ref) if silu else ref
    assert rel(y.permute(0, 3, 1, 2), ref) < 6e-4


@pytest.mark.parametrize("rows,C", [(4096, 320), (1000, 640), (512, 1280), (77, 768)])
def test_layer_norm(rows, C):
    x = h16(rnd(rows, C, seed=44) * 3 + 1)
    gamma, beta = rnd(C, seed=45) * 0.2 + 1, rnd(C, seed=46) * 0.2
    y = torch.empty_like(x)
    K.run_op(K.layer_norm_op(x, gamma, beta, y))
    assert rel(y, F.layer_norm(x.float(), (C,), gamma, beta, 1e-5)) < 6e-4


# ------------------------------------------------------------------------------- attention
@pytest.mark.parametrize("B,Nq,Nk,heads,d", [(2, 4096, 4096, 8, 40), (2, 1024, 1024, 8, 80), (2, 256, 256, 8, 160),
                                             (3, 64, 64, 8, 160), (2, 4096, 77, 8, 40), (2, 1024, 77, 8, 80),
                                             (2, 256, 77, 8, 160), (1, 200, 333, 4, 64), (1, 128, 64, 2, 128)])
def test_attention(B, Nq, Nk, heads, d):
    q = h16(rnd(B, Nq, heads * d, seed=50))
    k = h16(rnd(B, Nk, heads * d, seed=51))
    v = h16(rnd(B, Nk, heads * d, seed=52))
    o = torch.empty_like(q)
    K.run_op(K.attention_op(q, k, v, o, heads))
    qf, kf, vf = (t.float().view(B, -1, heads, d).transpose(1, 2) for t in (q, k, v))
    ref = F.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(B, Nq, heads * d)
    assert rel(o, ref) < 2e-3, rel(o, ref)


@pytest.mark.parametrize("B,Nq,Nk,heads,d", [(2, 4096, 77, 8, 40), (3, 1024, 77, 8, 80), (1, 130, 77, 2, 40), (5, 256, 80, 3, 80),
                                             (2, 1000, 1, 4, 40), (1, 64, 13, 8, 40), (150, 128, 77, 1, 40)])
def test_cross_attention_short_context(B, Nq, Nk, heads, d):
    """The dedicated short-context kernel (csrc/xattn.cu: Nk <= 80, d in {40, 80}): one 80-key tile, persistent CTAs
    walking (batch, query tile, head) units; ragged Nq, Nk below / at the tile size, odd unit counts, peaked scores."""
    q = h16(rnd(B, Nq, heads * d, seed=56) * 2.0)
    k = h16(rnd(B, Nk, heads * d, seed=57) * 2.0)
    v = h16(rnd(B, Nk, heads * d, seed=58))
    o = torch.full_like(q, 9.0)
    K.run_op(K.attention_op(q, k, v, o, heads))
    qf, kf, vf = (t.float().view(B, -1, heads, d).transpose(1, 2) for t in (q, k, v))
    ref = F.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(B, Nq, heads * d)
    assert rel(o, ref) < 2e-3, rel(o, ref)
    # strided views: q / o as channel slices of wider buffers, k / v as slices of one context projection
    wide = h16(rnd(B, Nq, heads * d + 64, seed=59))
    kv = h16(rnd(B, Nk, 2 * heads * d, seed=60))
    ow = torch.zeros(B, Nq, heads * d + 32, dtype=torch.float16, device=DEV)
    K.run_op(K.attention_op(wide[:, :, 64:], kv[:, :, :heads * d], kv[:, :, heads * d:], ow[:, :, 32:], heads))
    qf, kf, vf = (t.float().reshape(B, -1, heads, d).transpose(1, 2) for t in (wide[:, :, 64:], kv[:, :, :heads * d], kv[:, :, heads * d:]))
    ref = F.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(B, Nq, heads * d)
    assert rel(ow[:, :, 32:], ref) < 2e-3
    assert float(ow[:, :, :32].abs().max()) == 0.0


def test_cross_attention_is_repeatable():
    """Three warpgroups, two rings and split MMA issuers hand buffers around through mbarriers: 40 launches of the SD1.5
    shape must be bit-identical."""
    B, Nq, Nk, heads, d = 4, 4096, 77, 8, 40
    q, k, v = h16(rnd(B, Nq, heads * d, seed=61)), h16(rnd(B, Nk, heads * d, seed=62)), h16(rnd(B, Nk, heads * d, seed=63))
    o = torch.zeros_like(q)
    prog = K.Program()
    prog.add(K.attention_op(q, k, v, o, heads))
    prog.run()
    first = o.clone()
    for _ in range(40):
        o.zero_()
        prog.run()
        assert torch.equal(o, first)


def test_attention_peaked_scores_rescale_path():
    # large-magnitude q/k make the running max jump by more than the lazy-rescale threshold
    B, N, heads, d = 1, 512, 2, 40
    q, k = h16(rnd(B, N, heads * d, seed=53) * 4), h16(rnd(B, N, heads * d, seed=54) * 4)
    v = h16(rnd(B, N, heads * d, seed=55))
    o = torch.empty_like(q)
    K.run_op(K.attention_op(q, k, v, o, heads))
    qf, kf, vf = (t.float().view(B, -1, heads, d).transpose(1, 2) for t in (q, k, v))
    ref = F.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(B, N, heads * d)
    assert rel(o, ref) < 2e-3


@pytest.mark.parametrize("B,Nq,Nk,heads,d", [(1, 700, 1000, 2, 40), (2, 513, 257, 3, 40), (1, 300, 640, 2, 80),
                                             (1, 128, 320, 1, 160), (1, 1024, 2048, 2, 64)])
def test_attention_growing_row_max(B, Nq, Nk, heads, d):
    """Scores whose row maximum keeps growing along the key axis: every tile takes the lazy O-rescale path of the
    online softmax (ragged Nq / Nk, partially filled query groups)."""
    g = torch.Generator().manual_seed(170 + Nq)
    q = torch.randn(B, Nq, heads * d, generator=g).to(DEV)
    k = torch.randn(B, Nk, heads * d, generator=g).to(DEV)
    ramp = torch.linspace(0.2, 6.0, Nk, device=DEV).view(1, Nk, 1)        # later keys dominate: max grows by >> 2^8
    k = h16(k.abs() * ramp * q.sign().mean(1, keepdim=True).sign())
    q, v = h16(q.abs() * q.sign().mean(1, keepdim=True).sign()), h16(torch.randn(B, Nk, heads * d, generator=g).to(DEV))
    o = torch.zeros_like(q)
    K.run_op(K.attention_op(q, k, v, o, heads))
    qf, kf, vf = (t.float().view(B, -1, heads, d).transpose(1, 2) for t in (q, k, v))
    s = (qf @ kf.transpose(-1, -2)) / d ** 0.5
    assert float((s.cummax(-1).values[..., -1] - s.cummax(-1).values[..., 63]).min()) * 1.4427 > 8.0   # the path is taken
    ref = F.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(B, Nq, heads * d)
    assert torch.isfinite(o).all()
    assert rel(o, ref) < 3e-3, rel(o, ref)


@pytest.mark.parametrize("B,N,heads", [(1, 77, 12), (3, 77, 12), (2, 64, 4), (2, 130, 2), (1, 256, 1)])
def test_attention_causal(B, N, heads):
    """The CLIP text tower's attention (LD.py:4443-4449, 4291): -inf above the diagonal, d = 64."""
    d = 64
    q, k, v = (h16(rnd(B, N, heads * d, seed=140 + i) * (2.0 if i < 2 else 1.0)) for i in range(3))
    o = torch.zeros_like(q)
    L.check(L.lib().ldx_attention_causal(q.data_ptr(), heads * d, k.data_ptr(), heads * d, v.data_ptr(), heads * d,
                                         o.data_ptr(), heads * d, B, N, heads, d, 0))
    qf, kf, vf = (t.float().view(B, N, heads, d).transpose(1, 2) for t in (q, k, v))
    mask = torch.full((N, N), float("-inf"), device=DEV).triu_(1)
    ref = F.scaled_dot_product_attention(qf, kf, vf, attn_mask=mask).transpose(1, 2).reshape(B, N, heads * d)
    assert rel(o, ref) < 2e-3
    # the first token attends to itself only
    assert rel(o[:, 0], v[:, 0].float()) < 1e-3
    # strided q/k/v (views of one fused projection), through a recorded program
    qkv = h16(rnd(B, N, 3 * heads * d, seed=150))
    o2 = torch.zeros(B, N, heads * d, dtype=torch.float16, device=DEV)
    c = heads * d
    K.run_op(K.attention_op(qkv[:, :, :c], qkv[:, :, c:2 * c], qkv[:, :, 2 * c:], o2, heads, causal=True))
    q2, k2, v2 = (qkv[:, :, i * c:(i + 1) * c].float().reshape(B, N, heads, d).transpose(1, 2) for i in range(3))
    ref2 = F.scaled_dot_product_attention(q2, k2, v2, attn_mask=mask).transpose(1, 2).reshape(B, N, c)
    assert rel(o2, ref2) < 2e-3
    assert L.lib().ldx_attention_causal(q.data_ptr(), heads * d, k.data_ptr(), heads * d, v.data_ptr(), heads * d,
                                        o.data_ptr(), heads * d, B, N, heads, 40, 0) != 0     # d = 64 only


def test_linear_quick_gelu_and_clip_embed():
    M, N, Kd = 154, 3072, 768
    a, w, bias = h16(rnd(M, Kd, seed=160)), h16(rnd(N, Kd, seed=161, scale=Kd ** -0.5)), rnd(N, seed=162)
    out = torch.empty(M, N, dtype=torch.float16, device=DEV)
    K.run_gemm(K.linear_desc(a, K.pack_linear(w), out, bias=bias, flags=L.EPI_QUICK_GELU))
    x = a.float() @ w.float().t() + bias
    assert rel(out, x * torch.sigmoid(1.702 * x)) < 1e-3
    T, C, vocab = 77, 768, 1000
    tok = torch.randint(0, vocab, (2 * T,), device=DEV, dtype=torch.int32)
    te, pe = rnd(vocab, C, seed=163), rnd(T, C, seed=164)
    emb = torch.empty(2 * T, C, dtype=torch.float16, device=DEV)
    L.check(L.lib().ldx_clip_embed(tok.data_ptr(), te.data_ptr(), pe.data_ptr(), emb.data_ptr(), C, 2 * T, T, C, 0))
    assert torch.equal(emb, (te[tok.long()] + pe.repeat(2, 1)).half())


# ------------------------------------------------------------------------------- small ops
def test_input_output_transforms():
    from oracle import ldm_oracle as O
    B, H, W = 3, 16, 24
    x, sigma = rnd(B, 4, H, W, seed=60) * 5, torch.tensor([14.6, 1.0, 0.03], device=DEV)
    y = torch.empty(B, H, W, 8, dtype=torch.float16, device=DEV)
    L.check(L.lib().ldx_nchw_to_nhwc_scaled(x.data_ptr(), sigma.data_ptr(), y.data_ptr(), B, 4, H, W, 8, 0))
    ref = (x / (sigma.view(-1, 1, 1, 1) ** 2 + 1) ** 0.5).to(torch.float16).permute(0, 2, 3, 1)
    assert torch.equal(y[..., :4], ref) and float(y[..., 4:].abs().max()) == 0
    # timestep embedding incl. the nearest-sigma lookup
    table = O.make_sigmas_table()
    sig = torch.cat([O.calculate_sigmas("karras", 20)[:-1], torch.tensor([0.5, 7.7])]).to(DEV)
    emb = torch.empty(len(sig), 320, dtype=torch.float16, device=DEV)
    L.check(L.lib().ldx_timestep_embedding(sig.data_ptr(), table.log_sigmas.to(DEV).data_ptr(), 1000, emb.data_ptr(),
                                           len(sig), 320, 0))
    t = O.timestep_of_sigma(sig.cpu(), table.log_sigmas).float()
    ref = O.timestep_embedding(t, 320).to(torch.float16).to(DEV)
    assert float((emb.float() - ref.float()).abs().max()) < 2e-3
    # eps -> denoised and the fused CFG variant
    eps = h16(rnd(2 * B, H, W, 8, seed=61))
    den = torch.empty(2 * B, 4, H, W, device=DEV)
    x2, s2 = torch.cat([x, x]), torch.cat([sigma, sigma])
    L.check(L.lib().ldx_eps_to_denoised(eps.data_ptr(), 8, x2.data_ptr(), s2.data_ptr(), den.data_ptr(), 2 * B, 4, H, W, 0))
    ref = x2 - eps[..., :4].float().permute(0, 3, 1, 2) * s2.view(-1, 1, 1, 1)
    assert torch.allclose(den, ref, rtol=1e-6, atol=1e-6)
    out = torch.empty(B, 4, H, W, device=DEV)
    L.check(L.lib().ldx_cfg_denoised(eps.data_ptr(), 8, x.data_ptr(), sigma.data_ptr(), 7.0, out.data_ptr(), B, 4, H, W, 0))
    assert torch.allclose(out, O.cfg_combine(ref[B:], ref[:B], 7.0), rtol=1e-5, atol=1e-5)
    out2 = torch.empty_like(out)
    L.check(L.lib().ldx_cfg_combine(ref[B:].contiguous().data_ptr(), ref[:B].contiguous().data_ptr(), 7.0,
                                    out2.data_ptr(), out2.numel(), 0))
    assert torch.allclose(out2, O.cfg_combine(ref[B:], ref[:B], 7.0), rtol=1e-6, atol=1e-6)


def test_vae_post_and_softmax():
    dec = h16(rnd(2 * 32 * 32, 8, seed=62))
    img = torch.empty(2 * 32 * 32, 3, device=DEV)
    L.check(L.lib().ldx_vae_postprocess(dec.data_ptr(), 8, img.data_ptr(), dec.shape[0], 0))
    assert torch.equal(img, torch.clamp((dec[:, :3].float() + 1.0) / 2.0, 0.0, 1.0))
    # 1024 .. 16384 take the register-resident kernel, 1000 / 4100 the generic one; ld > n checks the row strides
    for rows, n, ld in ((300, 4096, 4096), (37, 16384, 16384), (50, 1024, 2048), (20, 8192, 8192), (33, 1000, 1000),
                        (9, 4100, 4104)):
        s = rnd(rows, ld, seed=63 + rows) * 20
        s[:, 3] += 40.0                      # a dominant score per row
        ref = torch.softmax(s[:, :n] * 0.0442, -1)
        pm = torch.zeros(rows, ld, dtype=torch.float16, device=DEV)
        L.check(L.lib().ldx_softmax_rows(s.data_ptr(), ld, pm.data_ptr(), ld, rows, n, 0.0442, 0))
        assert rel(pm[:, :n], ref) < 1e-3, (rows, n)
        assert (pm[:, :n].float().sum(-1) - 1).abs().max() < 2e-3
        assert not pm[:, n:].any()


def test_vae_preprocess_and_posterior_sample():
    from oracle import ldm_oracle as O
    for npix in (2 * 24 * 40, 1024 * 1024 + 3):       # the second exceeds one wave of the capped grid
        pix = torch.rand(npix, 3, device=DEV)
        out = torch.full((npix, 8), 7.0, dtype=torch.float16, device=DEV)
        L.check(L.lib().ldx_vae_preprocess(pix.data_ptr(), out.data_ptr(), npix, 0))
        assert torch.equal(out[:, :3], (pix * 2.0 - 1.0).half())
        assert not out[:, 3:].any()
    assert L.lib().ldx_vae_preprocess(None, out.data_ptr(), 4, 0) != 0
    for b, zc, h, w in ((3, 4, 5, 7), (2, 4, 256, 320)):
        mom = rnd(b, h, w, 2 * zc, seed=64) * 3
        mom[0, 0, 0, zc:] = torch.tensor([-50.0, 50.0, 0.0, 1.0], device=DEV)     # exercises the logvar clamp
        noise = rnd(b, zc, h, w, seed=65)
        z = torch.empty(b, zc, h, w, device=DEV)
        L.check(L.lib().ldx_vae_sample_posterior(mom.data_ptr(), noise.data_ptr(), z.data_ptr(), b, zc, h * w, 0))
        ref = O.posterior_sample(mom.permute(0, 3, 1, 2), noise)
        assert torch.allclose(z, ref, rtol=1e-5, atol=1e-6)
        L.check(L.lib().ldx_vae_sample_posterior(mom.data_ptr(), None, z.data_ptr(), b, zc, h * w, 0))
        assert torch.equal(z, mom.permute(0, 3, 1, 2)[:, :zc])


def test_bislerp_matches_reference_golden(golden):
    """ldx_bislerp_axis through the reference-signature wrapper, against the reference's outputs (LD.py:429)."""
    from lightdiffusion_b200 import upscale as U
    g = golden("bislerp.pt")
    for name, c in g["cases"].items():
        x = c.get("x", g["x"]).to(DEV)
        out = U.common_upscale(x, c["width"], c["height"], "bislerp", "disabled")
        assert out.shape == c["out"].shape and out.dtype == x.dtype, name
        assert torch.allclose(out.cpu(), c["out"], rtol=1e-4, atol=1e-5), (name, (out.cpu() - c["out"]).abs().max())
    # hires-fix bridge shape: 64x64 -> 128x128 latents, against the restatement
    from oracle import ldm_oracle as O
    x = rnd(3, 4, 64, 64, seed=90)
    got, ref = U.bislerp(x, 128, 128), O.bislerp(x, 128, 128)
    bad = ((got - ref).abs() > 1e-5 + 1e-4 * ref.abs()).float().mean().item()
    print("bislerp 64->128: rel", rel(got, ref), "max abs", (got - ref).abs().max().item(), "outside tol", bad)
    # near-opposite neighbours (sin(omega) -> 0) amplify the last-ulp differences of acos/sin between CUDA and ATen
    assert rel(got, ref) < 1e-5 and bad < 1e-4
    assert U.bislerp(x.half(), 96, 80).dtype == torch.float16
    with pytest.raises(Exception):
        U.bislerp(x.cpu(), 128, 128)          # no CPU fallback
    with pytest.raises(Exception):
        U.bislerp(rnd(1, 17, 4, 4, seed=91), 8, 8)   # more than 16 channels


def test_sampler_step_kernels():
    n = 8 * 4 * 64 * 64
    x, d, od, nz = (rnd(n, seed=70 + i) for i in range(4))
    x1 = x.clone()
    L.check(L.lib().ldx_dpmpp2m_step(x1.data_ptr(), d.data_ptr(), None, 0.8, 0.2, 0.0, n, 0))
    assert torch.allclose(x1, 0.8 * x + 0.2 * d, rtol=1e-6, atol=1e-6)
    x2 = x.clone()
    L.check(L.lib().ldx_dpmpp2m_step(x2.data_ptr(), d.data_ptr(), od.data_ptr(), 0.8, 0.2, 0.05, n, 0))
    assert torch.allclose(x2, 0.8 * x + 0.2 * d + 0.05 * (d - od), rtol=1e-6, atol=1e-6)
    x3 = x.clone()
    L.check(L.lib().ldx_euler_ancestral_step(x3.data_ptr(), d.data_ptr(), nz.data_ptr(), 2.0, 1.5, 0.7, n, 0))
    assert torch.allclose(x3, x + (x - d) / 2.0 * (1.5 - 2.0) + nz * 0.7, rtol=1e-6, atol=1e-6)


def test_errors_are_loud():
    a = h16(rnd(64, 60, seed=80))          # K pitch 60 -> TMA stride not a multiple of 16 bytes
    w, out = h16(rnd(32, 64, seed=81)), torch.empty(64, 32, dtype=torch.float16, device=DEV)
    with pytest.raises(L.LdxError):
        K.run_gemm(K.linear_desc(a, w, out))
    q = h16(rnd(1, 64, 8 * 24, seed=82))   # head dim 24 unsupported
    with pytest.raises(L.LdxError):
        K.run_op(K.attention_op(q, q, q, torch.empty_like(q), 8))


@pytest.mark.parametrize("B,H,W,cin,cout,pitch", [(2, 40, 24, 128, 3, 128), (3, 64, 64, 320, 4, 320), (1, 8, 8, 64, 1, 64),
                                                   (2, 17, 29, 72, 5, 80), (1, 14, 14, 64, 3, 64), (1, 15, 43, 128, 2, 136)])
def test_conv3x3_small_matches_conv2d(B, H, W, cin, cout, pitch):
    """ldx_conv3x3_small (VAE conv_out 128 -> 3, UNet out 320 -> 4): every input pixel read once, 9 * Cout per-tap
    partials on the tensor core, taps gathered in the epilogue.  Tile edges (14 x 14 interiors), ragged sizes, a
    channel pitch wider than Cin and a channel count off the 64 grid are covered; both outputs are checked."""
    g = torch.Generator().manual_seed(77 + H)
    buf = torch.randn(B, H, W, pitch, generator=g).half().to(DEV)
    x = buf[..., :cin]
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)).half().to(DEV)
    bias = (0.1 * torch.randn(cout, generator=g)).to(DEV)
    out16 = torch.zeros(B, H, W, 8, dtype=torch.float16, device=DEV)
    img = torch.zeros(B, H, W, cout, dtype=torch.float32, device=DEV)
    K.run_op(K.conv3x3_small_op(x, K.pack_conv3x3_small(w), bias, cout, out16=out16, out_img=img, post=1))
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1).permute(0, 2, 3, 1)
    assert rel(out16[..., :cout], ref) < 1e-3
    assert float(out16[..., cout:].abs().max()) == 0.0                # channels past Cout are not touched
    assert torch.equal(img, ((out16[..., :cout].float() + 1) / 2).clamp(0, 1))
    img2 = torch.zeros_like(img)
    K.run_op(K.conv3x3_small_op(x, K.pack_conv3x3_small(w), None, cout, out_img=img2))       # no bias, raw fp32 output
    assert rel(img2, F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), None, padding=1).permute(0, 2, 3, 1)) < 1e-3


@pytest.mark.parametrize("M,N,Kd,res", [(4096, 1280, 1280, True), (4096, 1280, 5120, False), (2048 * 8, 640, 2560, False),
                                        (4096, 1000, 1280, True), (4096 + 128, 1280, 2560, True), (32768, 320, 2880, False)])
def test_linear_two_accumulator_tiles(M, N, Kd, res):
    """N tiles wider than one 256-column accumulator (256 x 320 ... 256 x 512 pair tiles = two tcgen05 accumulators
    sharing every A stage, single-buffered TMEM): chosen by the cost model where it saves a wave (N = 1280 at
    M = 4096: 64 tiles on 74 CTA pairs instead of 128).  Ragged N / M and the fused-residual identity steps included."""
    a, w = h16(rnd(M, Kd, seed=51)), h16(rnd(N, Kd, seed=52, scale=Kd ** -0.5))
    bias = rnd(N, seed=53)
    r = h16(rnd(M, N, seed=54)) if res else None
    out = torch.empty(M, N, dtype=torch.float16, device=DEV)
    d = K.linear_desc(a, w, out, bias=bias, residual=r)
    import os
    if os.environ.get("LDX_GEMM_DUAL", "1") != "0" and N % 320 == 0:
        n_tiles = K.rowstats_parts(d) // 2
        assert -(-N // n_tiles) > 256, (N, n_tiles)          # the wide-tile path is what this test runs
    K.run_gemm(d)
    ref = a.float() @ w.float().t() + bias + (r.float() if res else 0)
    assert rel(out, ref) < 1e-3, (M, N, Kd, rel(out, ref))
    st = K.attach_rowstats(d, M)                             # row statistics out of a wide tile (LayerNorm folding)
    K.run_gemm(d)
    s = st.sum(1)
    o32 = out.float()
    assert rel(s[:, 0], o32.sum(1)) < 1e-3 and rel(s[:, 1], (o32 * o32).sum(1)) < 1e-3


def test_conv3x3_two_accumulator_tiles():
    """3x3 conv 320 -> 320 on a 16 x 64 x 64 map (M = 65536, K = 2880) with the timestep row add: 256 x 320 tiles."""
    B, H, W, C = 16, 64, 64, 320
    x = h16(rnd(B, H, W, C, seed=61))
    w = h16(rnd(C, C, 3, 3, seed=62, scale=(9 * C) ** -0.5))
    bias, emb = rnd(C, seed=63), rnd(B, C, seed=64)
    out = torch.empty(B, H, W, C, dtype=torch.float16, device=DEV)
    K.run_gemm(K.conv3x3_desc(x, K.pack_conv3x3(w), out, bias=bias, rowadd=emb))
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1) + emb[:, :, None, None]
    assert rel(out.permute(0, 3, 1, 2), ref) < 1e-3


# ------------------------------------------------------------------------------- GroupNorm statistics from producer epilogues
def _cs_reference(out_bhwc):
    o = out_bhwc.float()
    return torch.stack([o.sum((1, 2)), (o * o).sum((1, 2))], -1)          # [B, C, 2]


@pytest.mark.parametrize("B,H,W,cin,cout", [(4, 32, 32, 320, 320), (2, 24, 40, 64, 640), (3, 16, 16, 128, 1280), (1, 64, 64, 320, 960)])
def test_gemm_colstats_conv(B, H, W, cin, cout):
    """Column partials written by the conv epilogue (per 32-row slot: sum and sum of squares of the fp16 outputs) add
    up to the per-(sample, channel) sums of the produced tensor; ragged tiles contribute zeros."""
    x = h16(rnd(B, H, W, cin, seed=71))
    w = h16(rnd(cout, cin, 3, 3, seed=72, scale=(9 * cin) ** -0.5))
    out = torch.empty(B, H, W, cout, dtype=torch.float16, device=DEV)
    wp, bias = K.pack_conv3x3(w), rnd(cout, seed=73)         # descriptors hold raw pointers: keep the tensors alive
    d = K.conv3x3_desc(x, wp, out, bias=bias)
    r = K.attach_colstats(d, B, cout)
    assert r is not None
    buf, slots = r
    buf.fill_(float("nan"))                                  # every slot must be written (no stale entries)
    K.run_gemm(d)
    got = buf.sum(1)
    ref = _cs_reference(out)
    assert torch.isfinite(got).all()
    assert rel(got[..., 0], ref[..., 0]) < 1e-5 and rel(got[..., 1], ref[..., 1]) < 1e-5


@pytest.mark.parametrize("silu", [False, True])
def test_group_norm_from_colstats_concat(silu):
    """GroupNorm(32) over the channel concat [x0 | x1] with the statistics reduced from the producers' partials:
    x0 comes out of a 3x3 conv (4-D tiles), x1 out of a [rows, C] linear with a fused residual (slots over all rows),
    the group boundaries straddle the concat point (C = 320 + 640 = 960, 30 channels per group)."""
    B, H, W, c0, c1 = 2, 32, 32, 320, 640
    xin = h16(rnd(B, H, W, 64, seed=81))
    w0 = h16(rnd(c0, 64, 3, 3, seed=82, scale=(9 * 64) ** -0.5))
    x0 = torch.empty(B, H, W, c0, dtype=torch.float16, device=DEV)
    wp0, b0 = K.pack_conv3x3(w0), rnd(c0, seed=83) + 0.5
    d0 = K.conv3x3_desc(xin, wp0, x0, bias=b0)
    cs0 = K.attach_colstats(d0, B, c0)
    K.run_gemm(d0)
    a = h16(rnd(B * H * W, 320, seed=84))
    w1 = h16(rnd(c1, 320, seed=85, scale=320 ** -0.5))
    res = h16(rnd(B * H * W, c1, seed=86))
    x1 = torch.empty(B, H, W, c1, dtype=torch.float16, device=DEV)
    d1 = K.linear_desc(a, w1, x1, residual=res)
    r1 = K.attach_colstats(d1, 1, c1)
    K.run_gemm(d1)
    slots1 = H * W // 32
    cs1 = (r1[0].view(-1)[:B * slots1 * c1 * 2].view(B, slots1, c1, 2), slots1)
    C = c0 + c1
    gamma, beta = rnd(C, seed=87) * 0.2 + 1, rnd(C, seed=88) * 0.2
    y = torch.empty(B, H, W, C, dtype=torch.float16, device=DEV)
    y2 = torch.empty_like(y)
    stats = torch.empty(B * L.GN_STATS_FLOATS_PER_BATCH, dtype=torch.float32, device=DEV)
    K.run_op(K.group_norm_op(x0, x1, gamma, beta, y, stats, 1e-5, silu, colstats=(cs0, cs1)))
    K.run_op(K.group_norm_op(x0, x1, gamma, beta, y2, stats, 1e-5, silu))                 # two-pass path
    xc = torch.cat([x0, x1], -1).float().permute(0, 3, 1, 2)
    ref = F.group_norm(xc, 32, gamma, beta, 1e-5)
    ref = (F.silu(ref) if silu else ref).permute(0, 2, 3, 1)
    assert rel(y, ref) < 6e-4 and rel(y2, ref) < 6e-4
    assert rel(y, y2) < 1e-4


def test_gemm_colstats_upsample_phases_share_a_buffer():
    """The four phase GEMMs of a folded nearest-2x + 3x3 conv write disjoint slot ranges of one partial buffer."""
    B, H, W, C = 2, 16, 16, 128
    x = h16(rnd(B, H, W, C, seed=91))
    w = h16(rnd(C, C, 3, 3, seed=92, scale=(9 * C) ** -0.5))
    out = torch.empty(B, 2 * H, 2 * W, C, dtype=torch.float16, device=DEV)
    wp, bias = K.pack_upconv3x3(w), rnd(C, seed=93)
    descs = K.upconv3x3_descs(x, wp, out, bias=bias)
    buf, slots = K.attach_colstats(descs, B, C)
    buf.fill_(float("nan"))
    for d in descs:
        K.run_gemm(d)
    got, ref = buf.sum(1), _cs_reference(out)
    assert rel(got[..., 0], ref[..., 0]) < 1e-5 and rel(got[..., 1], ref[..., 1]) < 1e-5


@pytest.mark.parametrize("M", [256, 4096 + 130, 37])
def test_ff_geglu_fused(M):
    """x + ff(norm3(x)) (LD.py:4155-4162, FeedForward / GEGLU LD.py:4508-4536) in one launch at C = 320: the GEGLU
    intermediate goes epilogue -> shared memory -> second GEMM; ragged M, several tiles per CTA pair."""
    C, Fd = 320, 1280
    x = h16(rnd(M, C, seed=200) * 1.2 + 0.3)
    gamma, beta = 1.0 + 0.2 * rnd(C, seed=201), 0.3 * rnd(C, seed=202)
    w1, b1 = h16(rnd(2 * Fd, C, seed=203, scale=C ** -0.5)), rnd(2 * Fd, seed=204, scale=0.3)
    w2, b2 = h16(rnd(C, Fd, seed=205, scale=Fd ** -0.5)), rnd(C, seed=206, scale=0.3)
    wf, gsum, bf = K.fold_layer_norm(w1, b1, gamma, beta)
    w1p, b1p = K.pack_geglu(wf, bf, 64)
    _, g1p = K.pack_geglu(wf, gsum, 64)
    xs = x.float()
    stats = torch.stack([xs.sum(1), (xs * xs).sum(1)], 1).view(M, 1, 2).contiguous()
    out = torch.full((M, C), 3.0, dtype=torch.float16, device=DEV)
    K.run_op(K.ff_fused_op(x, w1p, b1p, g1p, K.pack_linear(w2), b2, stats, out))
    y = F.layer_norm(xs, (C,), gamma, beta, 1e-5) @ w1.float().t() + b1
    g = h16(y[:, :Fd] * F.gelu(y[:, Fd:])).float()            # the reference's fp16 intermediate
    ref = xs + g @ w2.float().t() + b2
    assert rel(out, ref) < 1.5e-3, rel(out, ref)


def test_ff_geglu_fused_is_repeatable_and_row_local():
    """Four producer / issuer warps, eight epilogue warps and a CTA pair hand S, A2, O and three rings around through
    mbarriers: 30 launches at a multi-tile-per-pair size (ragged last tile) must be bit-identical, and a row's result must not
    depend on which tile / CTA pair computed it (the same rows placed at another offset of a larger batch)."""
    C, Fd, M = 320, 1280, 256 * 74 * 2 + 256 * 9 + 77
    x = h16(rnd(M, C, seed=210) * 1.1 - 0.2)
    gamma, beta = 1.0 + 0.2 * rnd(C, seed=211), 0.3 * rnd(C, seed=212)
    w1, b1 = h16(rnd(2 * Fd, C, seed=213, scale=C ** -0.5)), rnd(2 * Fd, seed=214, scale=0.3)
    w2, b2 = h16(rnd(C, Fd, seed=215, scale=Fd ** -0.5)), rnd(C, seed=216, scale=0.3)
    wf, gsum, bf = K.fold_layer_norm(w1, b1, gamma, beta)
    w1p, b1p = K.pack_geglu(wf, bf, 64)
    _, g1p = K.pack_geglu(wf, gsum, 64)
    w2p = K.pack_linear(w2)

    def run(xin):
        xs = xin.float()
        stats = torch.stack([xs.sum(1), (xs * xs).sum(1)], 1).view(xin.shape[0], 1, 2).contiguous()
        out = torch.zeros_like(xin)
        prog = K.Program()
        prog.add(K.ff_fused_op(xin, w1p, b1p, g1p, w2p, b2, stats, out))
        prog.run()
        return prog, out

    prog, out = run(x)
    first = out.clone()
    for _ in range(30):
        out.zero_()
        prog.run()
        assert torch.equal(out, first)
    # rows 1024 .. 1024 + 5000 of x as the head of their own (smaller) launch: other tiles, other CTA pairs, same bits
    _, sub = run(x[1024:1024 + 5000].contiguous())
    assert torch.equal(sub, first[1024:1024 + 5000])


@pytest.mark.parametrize("B,H,C", [(3, 32, 640), (2, 64, 320), (2, 16, 320)])
def test_group_norm_folded_into_proj_in_weights(B, H, C):
    """SpatialTransformer.norm -> proj_in (LD.py:4242-4250) with the GroupNorm folded into per-sample weights
    (ldx_gn_fold_weights + ldx_gemm_desc_t.w_batch_rows): the 1x1 conv reads the UN-normalised map.  The inputs carry a group
    mean several times their spread, the case in which rounding the folded weights would be amplified if the bias did not
    cancel it; row statistics of the output (for the LayerNorm folded into the next GEMM) are checked as well."""
    x = h16(rnd(B, H, H, C, seed=301) * 0.7 + rnd(B, 1, 1, 32, seed=302).repeat_interleave(C // 32, -1) * 3.0 + 1.5)
    gamma, beta = 1.0 + 0.3 * rnd(C, seed=303), 0.4 * rnd(C, seed=304)
    w, bias = h16(rnd(C, C, seed=305, scale=C ** -0.5)), 0.3 * rnd(C, seed=306)
    wp = K.pack_linear(w)
    coef = torch.empty(B, C, 2, dtype=torch.float32, device=DEV)
    stats = torch.empty(B * L.GN_STATS_FLOATS_PER_BATCH, dtype=torch.float32, device=DEV)
    tickets = torch.zeros(B, dtype=torch.int32, device=DEV)
    K.run_op(K.group_norm_coef_op(x, None, gamma, beta, coef, stats, tickets, 1e-6))
    wb = torch.empty(B, C, wp.stride(0), dtype=torch.float16, device=DEV)
    bb = torch.empty(B, C, dtype=torch.float32, device=DEV)
    K.run_op(K.gn_fold_weights_op(wp, bias, beta, coef, wb, bb))
    out = torch.full((B, H, H, C), 7.0, dtype=torch.float16, device=DEV)
    d = K.conv1x1_batched_desc(x, wb, bb, out)
    rs = K.attach_rowstats(d, B * H * H)
    K.run_gemm(d)
    xf = x.float().permute(0, 3, 1, 2)
    y = F.group_norm(xf, 32, gamma, beta, 1e-6).permute(0, 2, 3, 1)
    ref = y @ w.float().t() + bias
    assert rel(out, ref) < 1.5e-3, rel(out, ref)
    # the two-pass path (normalised map rounded to fp16, shared weights) for comparison: same error class
    y16 = torch.empty_like(x)
    K.run_op(K.group_norm_op(x, None, gamma, beta, y16, stats, 1e-6, False))
    out2 = torch.empty_like(out)
    K.run_gemm(K.linear_desc(y16, wp, out2, bias=bias))
    assert rel(out, ref) < 2.0 * rel(out2, ref) + 1e-4, (rel(out, ref), rel(out2, ref))
    got = rs.sum(1)
    o32 = out.float().reshape(-1, C)
    assert rel(got[:, 0], o32.sum(1)) < 2e-3 and rel(got[:, 1], (o32 * o32).sum(1)) < 2e-3
