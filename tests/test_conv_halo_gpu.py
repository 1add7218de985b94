"""Halo-tile 3x3 conv (csrc/conv_halo.cu) and the GroupNorm coefficients it consumes, against the torch ops they
replace (F.conv2d LD.py:2382, GroupNorm + SiLU LD.py:2390-2395 / 5220-5261) computed in fp32 from the same fp16 inputs,
and against the implicit-GEMM + two-pass GroupNorm path of the same library.  All calls go through the C ABI."""
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


def conv_ref(x_nhwc, w_oihw, bias=None):
    y = F.conv2d(x_nhwc.float().permute(0, 3, 1, 2), w_oihw.float(), bias, padding=1)
    return y.permute(0, 2, 3, 1)


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(1, 16, 8, 64, 64), (2, 16, 16, 64, 128), (3, 32, 32, 320, 320),
                                            (2, 64, 64, 8, 320), (2, 16, 16, 1280, 1280), (2, 32, 32, 640, 96),
                                            (1, 24, 40, 128, 320), (5, 16, 16, 192, 640), (1, 128, 128, 128, 128)])
def test_conv3x3_halo_plain(B, H, W, Cin, Cout):
    """No transform: the nine shifted descriptor views of one halo tile against F.conv2d (H = 24 leaves half a tile
    out of the image; B = 5 / odd tile counts leave a CTA of a pair without a tile)."""
    x = h16(rnd(B, H, W, Cin, seed=1))
    w = h16(rnd(Cout, Cin, 3, 3, seed=2, scale=(9 * Cin) ** -0.5))
    bias, temb = rnd(Cout, seed=3), rnd(B, Cout, seed=4)
    out = torch.full((B, H, W, Cout), 7.0, dtype=torch.float16, device=DEV)
    K.run_gemm(K.conv3x3_halo_desc([x], K.pack_conv3x3(w), out, bias=bias, rowadd=temb))
    ref = conv_ref(x, w, bias) + temb[:, None, None, :]
    assert rel(out, ref) < 1e-3, rel(out, ref)
    # the implicit-GEMM kernel computes the same sums in another order
    out2 = torch.empty_like(out)
    K.run_gemm(K.conv3x3_desc(x, K.pack_conv3x3(w), out2, bias=bias, rowadd=temb))
    assert rel(out, out2) < 5e-4


def test_conv3x3_halo_forced_tiles():
    """Explicit N tiles (narrow, odd multiples of 32, full width)."""
    B, H, W, Cin, Cout = 2, 32, 32, 128, 640
    x = h16(rnd(B, H, W, Cin, seed=5))
    w = h16(rnd(Cout, Cin, 3, 3, seed=6, scale=(9 * Cin) ** -0.5))
    ref = conv_ref(x, w)
    for bn in (64, 128, 160, 256):
        out = torch.zeros(B, H, W, Cout, dtype=torch.float16, device=DEV)
        K.run_gemm(K.conv3x3_halo_desc([x], K.pack_conv3x3(w), out, bn=bn))
        assert rel(out, ref) < 1e-3, (bn, rel(out, ref))


def _gn_ref(x, gamma, beta, eps, silu):
    y = F.group_norm(x.float().permute(0, 3, 1, 2), 32, gamma, beta, eps)
    if silu:
        y = F.silu(y)
    return y.permute(0, 2, 3, 1)


def _coef(x0, x1, gamma, beta, eps, colstats=None):
    b = x0.shape[0]
    c = x0.shape[-1] + (0 if x1 is None else x1.shape[-1])
    coef = torch.full((b, c, 2), float("nan"), dtype=torch.float32, device=DEV)
    stats = torch.empty(b * L.GN_STATS_FLOATS_PER_BATCH, dtype=torch.float32, device=DEV)
    tickets = torch.zeros(b, dtype=torch.int32, device=DEV)
    K.run_op(K.group_norm_coef_op(x0, x1, gamma, beta, coef, stats, tickets, eps, colstats=colstats))
    assert int(tickets.abs().sum()) == 0            # the finishing CTA leaves the tickets zeroed
    return coef


@pytest.mark.parametrize("B,H,W,C0,C1", [(2, 64, 64, 320, 0), (3, 32, 32, 640, 320), (2, 16, 16, 1280, 0),
                                         (2, 16, 16, 1280, 1280), (1, 128, 128, 128, 0), (4, 8, 8, 1280, 0)])
def test_group_norm_coef(B, H, W, C0, C1):
    """coef = (rstd * gamma, beta - mean * rstd * gamma): the two-pass path's statistics kernel with the last-CTA
    finish, and the one-CTA-per-group kernel of the small maps."""
    x0 = h16(rnd(B, H, W, C0, seed=10) * 1.5 + 0.4)
    x1 = h16(rnd(B, H, W, C1, seed=11) - 0.3) if C1 else None
    C = C0 + C1
    gamma, beta = 1.0 + 0.2 * rnd(C, seed=12), 0.3 * rnd(C, seed=13)
    coef = _coef(x0, x1, gamma, beta, 1e-5)
    xc = (x0 if x1 is None else torch.cat([x0, x1], -1)).double()
    xg = xc.reshape(B, H * W, 32, C // 32)
    mean = xg.mean((1, 3))
    var = xg.var((1, 3), unbiased=False)
    rstd = (var + 1e-5).rsqrt()
    sc = rstd.repeat_interleave(C // 32, 1) * gamma.double()
    sh = beta.double() - mean.repeat_interleave(C // 32, 1) * sc
    assert rel(coef[..., 0], sc) < 1e-5
    assert float((coef[..., 1].double() - sh).abs().max()) < 1e-4
    # and it reproduces the library's own GroupNorm bit for bit when applied the same way
    y = torch.empty(B, H, W, C, dtype=torch.float16, device=DEV)
    stats = torch.empty(B * L.GN_STATS_FLOATS_PER_BATCH, dtype=torch.float32, device=DEV)
    K.run_op(K.group_norm_op(x0, x1, gamma, beta, y, stats, 1e-5, False))
    mine = torch.addcmul(coef[:, None, None, :, 1], xc.float(), coef[:, None, None, :, 0])   # not fused: compare loosely
    assert rel(mine, y) < 1e-3


@pytest.mark.parametrize("B,H,W,Cin,Cout,silu,eps", [(2, 64, 64, 320, 320, True, 1e-5), (2, 32, 32, 640, 640, True, 1e-5),
                                                     (2, 16, 16, 1280, 1280, True, 1e-5), (1, 128, 128, 128, 128, True, 1e-6),
                                                     (3, 32, 32, 320, 640, False, 1e-6), (1, 24, 16, 64, 64, True, 1e-5)])
def test_conv3x3_halo_group_norm(B, H, W, Cin, Cout, silu, eps):
    """conv(silu(GroupNorm(x))) with the normalisation applied to the landed halo tile."""
    x = h16(rnd(B, H, W, Cin, seed=20) * 1.3 + 0.5)
    gamma, beta = 1.0 + 0.2 * rnd(Cin, seed=21), 0.3 * rnd(Cin, seed=22)
    w = h16(rnd(Cout, Cin, 3, 3, seed=23, scale=(9 * Cin) ** -0.5))
    bias = rnd(Cout, seed=24)
    coef = _coef(x, None, gamma, beta, eps)
    out = torch.zeros(B, H, W, Cout, dtype=torch.float16, device=DEV)
    K.run_gemm(K.conv3x3_halo_desc([x], K.pack_conv3x3(w), out, gn=(coef, silu), bias=bias))
    ref = conv_ref(h16(_gn_ref(x, gamma, beta, eps, silu)), w, bias)
    assert rel(out, ref) < 1.5e-3, rel(out, ref)
    # the same library's two-pass GroupNorm + implicit-GEMM conv: identical normalised operand, other summation order
    y = torch.empty_like(x)
    stats = torch.empty(B * L.GN_STATS_FLOATS_PER_BATCH, dtype=torch.float32, device=DEV)
    K.run_op(K.group_norm_op(x, None, gamma, beta, y, stats, eps, silu))
    out2 = torch.empty_like(out)
    K.run_gemm(K.conv3x3_desc(y, K.pack_conv3x3(w), out2, bias=bias))
    assert rel(out, out2) < 5e-4, rel(out, out2)


def test_conv3x3_halo_concat_skip_residual():
    """The up-path ResBlock shapes: GroupNorm + conv over cat([h, hsp]) (two halo sources), then the second conv with
    the 1x1 skip_connection over the raw concat as extra K segments (LD.py:5267, 5749), and the identity-residual
    form of the equal-width blocks."""
    B, H, W, C0, C1, Cout = 2, 32, 32, 640, 320, 640
    h0, h1 = h16(rnd(B, H, W, C0, seed=30) + 0.2), h16(rnd(B, H, W, C1, seed=31) * 0.7)
    C = C0 + C1
    gamma, beta = 1.0 + 0.2 * rnd(C, seed=32), 0.3 * rnd(C, seed=33)
    w1 = h16(rnd(Cout, C, 3, 3, seed=34, scale=(9 * C) ** -0.5))
    coef = _coef(h0, h1, gamma, beta, 1e-5)
    t1 = torch.zeros(B, H, W, Cout, dtype=torch.float16, device=DEV)
    K.run_gemm(K.conv3x3_halo_desc([h0, h1], K.pack_conv3x3(w1), t1, gn=(coef, True)))
    xc = torch.cat([h0, h1], -1)
    ref1 = conv_ref(h16(_gn_ref(xc, gamma, beta, 1e-5, True)), w1)
    assert rel(t1, ref1) < 1.5e-3, rel(t1, ref1)

    # second conv: GroupNorm(t1) -> conv3x3, plus skip 1x1 over [h0 | h1]
    g2, b2 = 1.0 + 0.2 * rnd(Cout, seed=35), 0.3 * rnd(Cout, seed=36)
    w2 = h16(rnd(Cout, Cout, 3, 3, seed=37, scale=(9 * Cout) ** -0.5))
    wskip = h16(rnd(Cout, C, seed=38, scale=C ** -0.5))
    wcat = torch.cat([K.pack_conv3x3(w2), K.pack_linear(wskip)], 1).contiguous()
    coef2 = _coef(t1, None, g2, b2, 1e-5)
    out = torch.zeros(B, H, W, Cout, dtype=torch.float16, device=DEV)
    K.run_gemm(K.conv3x3_halo_desc([t1], wcat, out, gn=(coef2, True), extra_srcs=[h0, h1],
                                   extra_segs=[(0, 0, 0, 0, C0 // 64), (1, 0, 0, 0, C1 // 64)]))
    ref = conv_ref(h16(_gn_ref(t1, g2, b2, 1e-5, True)), w2) + xc.float() @ wskip.float().t()
    assert rel(out, ref) < 1.5e-3, rel(out, ref)

    # equal widths: x + conv(...) with the residual riding the tensor core
    res = h16(rnd(B, H, W, Cout, seed=39))
    out3 = torch.zeros(B, H, W, Cout, dtype=torch.float16, device=DEV)
    K.run_gemm(K.conv3x3_halo_desc([t1], K.pack_conv3x3(w2), out3, gn=(coef2, True), residual=res))
    ref3 = conv_ref(h16(_gn_ref(t1, g2, b2, 1e-5, True)), w2) + res.float()
    assert rel(out3, ref3) < 1.5e-3, rel(out3, ref3)


def test_conv3x3_halo_colstats_chain():
    """A halo conv that leaves GroupNorm column partials for the next normalisation (maps >= 128 x 128)."""
    B, H, W, C = 1, 128, 128, 128
    x = h16(rnd(B, H, W, C, seed=40))
    w = h16(rnd(C, C, 3, 3, seed=41, scale=(9 * C) ** -0.5))
    out = torch.zeros(B, H, W, C, dtype=torch.float16, device=DEV)
    d = K.conv3x3_halo_desc([x], K.pack_conv3x3(w), out)
    cs = K.attach_colstats(d, B, C)
    assert cs is not None
    K.run_gemm(d)
    buf, slots = cs
    o = out.float().reshape(B, H * W, C)
    assert rel(buf[..., 0].sum(1), o.sum(1)) < 1e-4
    assert rel(buf[..., 1].sum(1), (o * o).sum(1)) < 1e-4
    gamma, beta = 1.0 + 0.2 * rnd(C, seed=42), 0.3 * rnd(C, seed=43)
    c1 = _coef(out, None, gamma, beta, 1e-6)
    c2 = _coef(out, None, gamma, beta, 1e-6, colstats=((buf, slots), None))
    assert rel(c2, c1) < 1e-5


@pytest.mark.parametrize("B,H,W,C,N", [(2, 64, 64, 320, 320), (3, 32, 32, 640, 640), (2, 16, 16, 1280, 1280),
                                       (1, 24, 16, 64, 96)])
def test_conv1x1_group_norm_rowstats(B, H, W, C, N):
    """SpatialTransformer.norm + proj_in (LD.py:4246-4251): GroupNorm without SiLU folded into the operand path of a
    1x1 conv that also leaves the LayerNorm row statistics of its output."""
    x = h16(rnd(B, H, W, C, seed=60) * 1.2 - 0.4)
    gamma, beta = 1.0 + 0.2 * rnd(C, seed=61), 0.3 * rnd(C, seed=62)
    w, bias = h16(rnd(N, C, seed=63, scale=C ** -0.5)), rnd(N, seed=64)
    coef = _coef(x, None, gamma, beta, 1e-6)
    out = torch.zeros(B, H, W, N, dtype=torch.float16, device=DEV)
    d = K.conv1x1_gn_desc(x, K.pack_linear(w), out, gn=(coef, False), bias=bias)
    stats = K.attach_rowstats(d, B * H * W)
    K.run_gemm(d)
    ref = h16(_gn_ref(x, gamma, beta, 1e-6, False)).float() @ w.float().t() + bias
    assert rel(out, ref) < 1.5e-3, rel(out, ref)
    o = out.float().reshape(B * H * W, N)
    assert rel(stats[:, :, 0].sum(1), o.sum(1)) < 2e-3
    assert rel(stats[:, :, 1].sum(1), (o * o).sum(1)) < 2e-3


def test_halo_conv_is_repeatable():
    """The operand transform rewrites shared memory between the TMA landing and the MMA (generic-proxy stores, proxy
    fence, cross-CTA arrive): a missing fence or barrier shows up as run-to-run differences.  40 launches of the
    up-path shape (two sources, GroupNorm + SiLU, CTA pairs, several tiles per CTA) must be bit-identical."""
    B, H, W, C0, C1, Cout = 4, 64, 64, 320, 320, 320
    h0, h1 = h16(rnd(B, H, W, C0, seed=70) + 0.3), h16(rnd(B, H, W, C1, seed=71))
    gamma, beta = 1.0 + 0.2 * rnd(C0 + C1, seed=72), 0.3 * rnd(C0 + C1, seed=73)
    w = K.pack_conv3x3(h16(rnd(Cout, C0 + C1, 3, 3, seed=74, scale=(9 * (C0 + C1)) ** -0.5)))
    temb = rnd(B, Cout, seed=75)
    coef = _coef(h0, h1, gamma, beta, 1e-5)
    prog = K.Program()
    out = torch.zeros(B, H, W, Cout, dtype=torch.float16, device=DEV)
    prog.add(K.conv3x3_halo_desc([h0, h1], w, out, gn=(coef, True), rowadd=temb))
    prog.run()
    first = out.clone()
    for _ in range(40):
        out.zero_()
        prog.run()
        assert torch.equal(out, first)
    # the coefficients themselves: atomic ticket + last-CTA finish, fixed summation order
    for _ in range(10):
        assert torch.equal(_coef(h0, h1, gamma, beta, 1e-5), coef)


def test_halo_argument_errors():
    x = h16(rnd(1, 16, 12, 64, seed=50))             # W % 8 != 0
    w = K.pack_conv3x3(h16(rnd(64, 64, 3, 3, seed=51)))
    out = torch.zeros(1, 16, 12, 64, dtype=torch.float16, device=DEV)
    with pytest.raises(L.LdxError):
        K.run_gemm(K.conv3x3_halo_desc([x], w, out))
    x = h16(rnd(1, 16, 16, 64, seed=52))
    out32 = torch.zeros(1, 16, 16, 64, dtype=torch.float32, device=DEV)
    with pytest.raises(L.LdxError):
        K.run_gemm(K.conv3x3_halo_desc([x], w, out32, flags=L.EPI_OUT_F32))
