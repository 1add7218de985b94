"""One small invocation of the hot path on cuda:0, checked against the oracle (used by
__graft_entry__.smoke()).

Order matters for the driver's launch capture (it records the first launches of this process): phase 1 runs
the hot kernels directly -- implicit-GEMM conv, linear, GroupNorm, flash attention (self and 77-key cross), the
fused CFG + DPM++ 2M step tail -- each against the torch fp32 op it replaces; phase 2 checks the per-step UNet
eps of a reduced-width SD1.x UNet against the oracle at the stated tolerance (2e-3 relative L2); phase 3 runs the
whole pipeline (3 sampler steps + VAE decode) and checks the decoded image (>= 40 dB PSNR).  Weights are prepared
on the host so that the first launches of the process are libldx.so kernels, not ATen copies.
"""
import math

import torch

EPS_TOL = 2e-3        # north_star: per-step UNet epsilon, relative L2, fp16
PSNR_MIN = 40.0       # north_star: decoded image
KERNEL_TOL = 1.5e-3   # one fp16-output kernel against the fp32 op on the same fp16 inputs


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm())


def _kernels(dev):
    import torch.nn.functional as F
    from . import _lib as L
    from . import kernels as K
    g = torch.Generator().manual_seed(1)
    rnd = lambda *s: torch.randn(*s, generator=g)
    out = {}
    # 3x3 conv 64 -> 64 on a 2 x 32 x 32 map + bias (implicit GEMM, tcgen05)
    x, w, b = rnd(2, 32, 32, 64).half().to(dev), (rnd(64, 64, 3, 3) / 24).half().to(dev), rnd(64).to(dev)
    y = torch.empty(2, 32, 32, 64, dtype=torch.float16, device=dev)
    K.run_gemm(K.conv3x3_desc(x, K.pack_conv3x3(w), y, bias=b))
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=1).permute(0, 2, 3, 1)
    out["conv3x3"] = _rel(y, ref)
    # linear 256 -> 320 with a fused residual
    a, wl, r = rnd(512, 256).half().to(dev), (rnd(320, 256) / 16).half().to(dev), rnd(512, 320).half().to(dev)
    yl = torch.empty(512, 320, dtype=torch.float16, device=dev)
    K.run_gemm(K.linear_desc(a, K.pack_linear(wl), yl, residual=r))
    out["linear+res"] = _rel(yl, a.float() @ wl.float().t() + r.float())
    # GroupNorm(32) + SiLU
    gx = rnd(2, 16, 16, 128).half().to(dev)
    gam, bet = (1 + 0.1 * rnd(128)).to(dev), (0.1 * rnd(128)).to(dev)
    gy = torch.empty_like(gx)
    stats = torch.empty(2 * L.GN_STATS_FLOATS_PER_BATCH, dtype=torch.float32, device=dev)
    K.run_op(K.group_norm_op(gx, None, gam, bet, gy, stats, 1e-5, True))
    out["group_norm"] = _rel(gy, F.silu(F.group_norm(gx.float().permute(0, 3, 1, 2), 32, gam, bet, 1e-5)).permute(0, 2, 3, 1))
    # the ResBlock form: GroupNorm(32) + SiLU applied in the operand path of the halo-tile 3x3 conv (statistics-only
    # GroupNorm launch + ldx_conv_halo_kernel), against the torch ops on the same fp16 inputs
    hx = (rnd(2, 32, 32, 128) * 1.3 + 0.4).half().to(dev)
    hw, hb = (rnd(128, 128, 3, 3) / 34).half().to(dev), rnd(128).to(dev)
    coef = torch.empty(2, 128, 2, dtype=torch.float32, device=dev)
    tickets = torch.zeros(2, dtype=torch.int32, device=dev)
    K.run_op(K.group_norm_coef_op(hx, None, gam, bet, coef, stats, tickets, 1e-5))
    hy = torch.empty(2, 32, 32, 128, dtype=torch.float16, device=dev)
    K.run_gemm(K.conv3x3_halo_desc([hx], K.pack_conv3x3(hw), hy, gn=(coef, True), bias=hb))
    hn = F.silu(F.group_norm(hx.float().permute(0, 3, 1, 2), 32, gam, bet, 1e-5)).half().float()
    out["gn+conv_halo"] = _rel(hy, F.conv2d(hn, hw.float(), hb, padding=1).permute(0, 2, 3, 1))
    # attention: 1024-token self-attention (d = 40, 8 heads) and 77-key cross-attention
    for name, nk in (("self_attn", 1024), ("cross_attn", 77)):
        q, k, v = rnd(2, 1024, 320).half().to(dev), rnd(2, nk, 320).half().to(dev), rnd(2, nk, 320).half().to(dev)
        o = torch.empty_like(q)
        K.run_op(K.attention_op(q, k, v, o, 8))
        sp = lambda t: t.float().view(2, -1, 8, 40).transpose(1, 2)
        out[name] = _rel(o, F.scaled_dot_product_attention(sp(q), sp(k), sp(v)).transpose(1, 2).reshape(2, 1024, 320))
    # fused step tail: CFG on the denoised predictions + DPM++ 2M update + old <- denoised + next sigma
    nb, C, H, W = 2, 4, 16, 16
    eps = rnd(2 * nb, H, W, 8).half().to(dev)
    xs, old = rnd(nb, C, H, W).to(dev), rnd(nb, C, H, W).to(dev)
    sig_io = torch.zeros(nb, device=dev)
    x2, old2 = xs.clone(), old.clone()
    sigma, scale, cx, cd, co = 3.0, 7.0, 0.61, 0.39, 0.17
    L.check(L.lib().ldx_cfg_dpmpp2m_step(eps.data_ptr(), 8, x2.data_ptr(), old2.data_ptr(), None, sigma, scale, cx, cd, co,
                                         1, 0, 2.0, sig_io.data_ptr(), nb, C, H, W,
                                         torch.cuda.current_stream().cuda_stream), "ldx_cfg_dpmpp2m_step")
    e = eps[..., :C].float().permute(0, 3, 1, 2)
    du, dc = xs - e[:nb] * sigma, xs - e[nb:] * sigma
    den = du + (dc - du) * scale
    out["step_tail"] = max(_rel(x2, cx * xs + cd * den + co * (den - old)), _rel(old2, den))
    assert float(sig_io[0]) == 2.0 and float(sig_io[1]) == 2.0
    return out


def run() -> None:
    from . import sd15
    from .engine import UNetEngine, VAEDecoderEngine
    from .pipeline import Txt2ImgPipeline, calculate_sigmas
    from oracle import ldm_oracle as O          # checker only

    assert torch.cuda.is_available(), "smoke() needs cuda:0"
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = "cuda:0"
    torch.cuda.set_device(0)

    # ---- phase 1: the hot kernels, one launch each
    kerr = _kernels(dev)
    print("smoke kernels:", ", ".join(f"{k} {v:.1e}" for k, v in kerr.items()))
    assert max(kerr.values()) < KERNEL_TOL, kerr

    # ---- phase 2: per-step UNet eps vs the oracle (weights fp16-rounded on the host, packed by ldx_pack_*)
    cfg = sd15.UNetConfig(model_channels=128, channel_mult=(1, 2, 2, 2), num_heads=2, context_dim=256)
    usd_h = {k: v.half() for k, v in sd15.seeded_state_dict(sd15.unet_param_shapes(cfg), 1234, bias_std=0.05,
                                                            norm_jitter=0.1).items()}
    vsd_h = {k: v.half() for k, v in sd15.seeded_state_dict(sd15.vae_decoder_param_shapes(), 99, bias_std=0.05,
                                                            norm_jitter=0.1).items()}
    pipe = Txt2ImgPipeline(UNetEngine(usd_h, cfg, dev), VAEDecoderEngine(vsd_h, device=dev))
    g = torch.Generator().manual_seed(0)
    b, hw = 1, 16
    noise = torch.randn(b, 4, hw, hw, generator=g).to(dev)
    cond, uncond = torch.randn(b, 77, 256, generator=g).to(dev), torch.randn(b, 77, 256, generator=g).to(dev)
    usd = {k: v.float().to(dev) for k, v in usd_h.items()}          # the oracle's copy (same fp16-rounded values)
    vsd = {k: v.float().to(dev) for k, v in vsd_h.items()}
    table = O.make_sigmas_table()
    worst = 0.0
    for s in (14.6146, 1.99, 0.0485):
        sigma = torch.full((b,), s, device=dev)
        xs = noise * math.sqrt(1 + s * s)
        e = pipe.unet.eps(xs, sigma, cond)
        xc = (xs / (sigma.view(-1, 1, 1, 1) ** 2 + 1) ** 0.5).half().float()
        t = O.timestep_of_sigma(sigma.cpu(), table.log_sigmas).float().to(dev)
        with torch.inference_mode():
            ref = O.unet_forward(usd, cfg, xc, t, cond.half().float())
        worst = max(worst, _rel(e, ref))
    print(f"smoke unet: per-step eps rel-L2 {worst:.2e} (tolerance {EPS_TOL:.0e})")
    assert worst < EPS_TOL, worst

    # ---- phase 3: the pipeline (CFG plan, fused step tail, VAE decode)
    sig = calculate_sigmas("karras", 3)
    z, img = pipe.generate(noise, cond, uncond, 7.0, sig, "dpmpp_2m")
    with torch.inference_mode():
        zr, imgr = O.txt2img(usd, cfg, vsd, sd15.SD15_VAE, noise, cond.half().float(), uncond.half().float(), 7.0,
                             sig.to(dev), "dpmpp_2m")
    rel = _rel(z, zr)
    mse = float(((img - imgr) ** 2).mean())
    psnr = 10 * math.log10(1.0 / max(mse, 1e-20))
    print(f"smoke: latent rel-L2 {rel:.2e}, image PSNR {psnr:.1f} dB, launches {pipe.gpu_launches}")
    # three guided steps compound the per-step error (each <= 2e-3, CFG scale 7): the trajectory bound is 1e-2
    assert rel < 1e-2 and psnr >= PSNR_MIN, (rel, psnr)
