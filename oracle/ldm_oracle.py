"""ORACLE — test infrastructure, not product code.

Plain-PyTorch fp32 restatement of the reference's SD1.5 denoising hot path
(/root/reference/LightDiffusion.py, cited below as LD.py:line).  It exists only to
check the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  The product package
(``lightdiffusion_b200``) never does.

Pinning: the reference has no tests and no golden vectors, so
*the reference's own tests pin nothing for this path*.  This restatement is pinned
against outputs of the reference code itself, executed in the build container by
``oracle/make_golden.py`` and committed under ``tests/golden/``; ``tests/test_oracle.py``
replays them.  Every function is state-dict driven (reference parameter names), so the
same seeded weights feed the reference, this oracle and the CUDA engine.

All arithmetic runs in the dtype of the inputs (fp32 in tests; TF32 must be disabled
when run on a GPU).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional

import torch
import torch.nn.functional as F

from lightdiffusion_b200.sd15 import (AttnSpec, ConvSpec, ResSpec, UNetConfig, VAEConfig,
                                      unet_layout, vae_decoder_levels, vae_encoder_levels)

SD = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------- schedules
class SigmaTable:
    """The 1000 discrete sigmas and their logs. LD.py:787-796 (scaled-linear betas in float64),
    LD.py:1314-1326; note log_sigmas is log(float64 sigmas) cast to fp32, not log(fp32 sigmas)."""

    def __init__(self, linear_start=0.00085, linear_end=0.012, timesteps=1000):
        betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=torch.float64) ** 2
        alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        sig64 = ((1 - alphas_cumprod) / alphas_cumprod) ** 0.5
        self.sigmas = sig64.float()
        self.log_sigmas = sig64.log().float()

    @property
    def sigma_min(self):
        return self.sigmas[0]

    @property
    def sigma_max(self):
        return self.sigmas[-1]


def make_sigmas_table() -> SigmaTable:
    return SigmaTable()


def timestep_of_sigma(sigma: torch.Tensor, log_sigmas: torch.Tensor) -> torch.Tensor:
    """Nearest-neighbour timestep. LD.py:1336-1339."""
    dists = sigma.log().to(log_sigmas.device) - log_sigmas[:, None]
    return dists.abs().argmin(dim=0).view(sigma.shape).to(sigma.device)


def sigma_of_timestep(t: torch.Tensor, log_sigmas: torch.Tensor) -> torch.Tensor:
    """Log-linear interpolation. LD.py:1341-1351."""
    t = torch.clamp(t.float(), min=0, max=len(log_sigmas) - 1)
    lo, hi, w = t.floor().long(), t.ceil().long(), t.frac()
    return ((1 - w) * log_sigmas[lo] + w * log_sigmas[hi]).exp()


def sigmas_karras(n: int, sigma_min: float, sigma_max: float, rho: float = 7.0) -> torch.Tensor:
    """LD.py:831-837."""
    ramp = torch.linspace(0, 1, n)
    lo, hi = sigma_min ** (1 / rho), sigma_max ** (1 / rho)
    s = (hi + ramp * (lo - hi)) ** rho
    return torch.cat([s, s.new_zeros([1])])


def sigmas_normal(n: int, table: SigmaTable) -> torch.Tensor:
    """LD.py:2639-2651."""
    ls = table.log_sigmas
    start = timestep_of_sigma(table.sigma_max, ls)
    end = timestep_of_sigma(table.sigma_min, ls)
    ts = torch.linspace(start, end, n)
    sigs = [float(sigma_of_timestep(t, ls)) for t in ts] + [0.0]
    return torch.FloatTensor(sigs)


def calculate_sigmas(name: str, steps: int, table: Optional[SigmaTable] = None) -> torch.Tensor:
    """LD.py:3045-3054."""
    table = make_sigmas_table() if table is None else table
    if name == "karras":
        return sigmas_karras(steps, float(table.sigma_min), float(table.sigma_max))
    if name == "normal":
        return sigmas_normal(steps, table)
    raise ValueError(name)


def sigmas_for_denoise(name: str, steps: int, denoise: Optional[float], table=None) -> torch.Tensor:
    """KSampler1.set_steps, LD.py:3097-3104."""
    if denoise is None or denoise > 0.9999:
        return calculate_sigmas(name, steps, table)
    new_steps = int(steps / denoise)
    return calculate_sigmas(name, new_steps, table)[-(steps + 1):]


# ----------------------------------------------------------------------------- UNet parts
def timestep_embedding(t: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    """LD.py:803-812 (cos first, then sin; fp32)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _lin(sd: SD, p: str, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _conv(sd: SD, p: str, x, stride=1, padding=0):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=padding)


def _gn(sd: SD, p: str, x, eps):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps)


def _ln(sd: SD, p: str, x):
    return F.layer_norm(x, x.shape[-1:], sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def attention(q, k, v, heads: int):
    """softmax(q k^T / sqrt(d)) v per head, no mask. LD.py:3966-3978."""
    b, nq, c = q.shape
    d = c // heads
    q, k, v = (t.view(b, -1, heads, d).transpose(1, 2) for t in (q, k, v))
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(d)
    o = torch.matmul(torch.softmax(s, dim=-1), v)
    return o.transpose(1, 2).reshape(b, nq, c)


def resblock(sd: SD, p: str, x, emb):
    """ResBlock1._forward, LD.py:5276-5287 (GN eps 1e-5)."""
    h = _conv(sd, p + ".in_layers.2", F.silu(_gn(sd, p + ".in_layers.0", x, 1e-5)), padding=1)
    h = h + _lin(sd, p + ".emb_layers.1", F.silu(emb))[:, :, None, None]
    h = _conv(sd, p + ".out_layers.3", F.silu(_gn(sd, p + ".out_layers.0", h, 1e-5)), padding=1)
    skip = _conv(sd, p + ".skip_connection", x) if (p + ".skip_connection.weight") in sd else x
    return skip + h


def cross_attention(sd: SD, p: str, x, context, heads):
    """CrossAttention.forward, LD.py:4029-4036."""
    ctx = x if context is None else context
    q, k, v = _lin(sd, p + ".to_q", x), _lin(sd, p + ".to_k", ctx), _lin(sd, p + ".to_v", ctx)
    return _lin(sd, p + ".to_out.0", attention(q, k, v, heads))


def transformer_block(sd: SD, p: str, x, context, heads):
    """BasicTransformerBlock._forward, LD.py:4117-4162; GEGLU LD.py:4508-4515 (erf GELU)."""
    x = x + cross_attention(sd, p + ".attn1", _ln(sd, p + ".norm1", x), None, heads)
    x = x + cross_attention(sd, p + ".attn2", _ln(sd, p + ".norm2", x), context, heads)
    a, gate = _lin(sd, p + ".ff.net.0.proj", _ln(sd, p + ".norm3", x)).chunk(2, dim=-1)
    return _lin(sd, p + ".ff.net.2", a * F.gelu(gate)) + x


def spatial_transformer(sd: SD, s: AttnSpec, x, context):
    """SpatialTransformer.forward, LD.py:4242-4262 (GN eps 1e-6, 1x1 conv projections)."""
    b, c, h, w = x.shape
    y = _conv(sd, s.name + ".proj_in", _gn(sd, s.name + ".norm", x, 1e-6))
    y = y.flatten(2).transpose(1, 2)
    y = transformer_block(sd, s.name + ".transformer_blocks.0", y, context, s.heads)
    y = y.transpose(1, 2).reshape(b, -1, h, w)
    return _conv(sd, s.name + ".proj_out", y) + x


def _run_block(sd, layers, h, emb, context, out_hw=None):
    for s in layers:
        if isinstance(s, ResSpec):
            h = resblock(sd, s.name, h, emb)
        elif isinstance(s, AttnSpec):
            h = spatial_transformer(sd, s, h, context)
        elif s.kind == "conv_in":
            h = _conv(sd, s.name, h, padding=1)
        elif s.kind == "down":
            h = _conv(sd, s.name, h, stride=2, padding=1)      # Downsample1, LD.py:5184
        else:                                                  # Upsample1, LD.py:5142-5152
            size = [h.shape[2] * 2, h.shape[3] * 2] if out_hw is None else list(out_hw)
            h = _conv(sd, s.name, F.interpolate(h, size=size, mode="nearest"), padding=1)
    return h


def unet_forward(sd: SD, cfg: UNetConfig, x, timesteps, context):
    """UNetModel1.forward, LD.py:5688-5767. x [B,4,H,W], timesteps [B], context [B,77,ctx]."""
    lay = unet_layout(cfg)
    t_emb = timestep_embedding(timesteps, cfg.model_channels).to(x.dtype)
    emb = _lin(sd, "time_embed.2", F.silu(_lin(sd, "time_embed.0", t_emb)))
    hs, h = [], x
    for layers in lay.input_blocks:
        h = _run_block(sd, layers, h, emb, context)
        hs.append(h)
    h = _run_block(sd, lay.middle, h, emb, context)
    for layers in lay.output_blocks:
        h = torch.cat([h, hs.pop()], dim=1)
        out_hw = hs[-1].shape[2:] if hs else None
        h = _run_block(sd, layers, h, emb, context, out_hw)
    return _conv(sd, "out.2", F.silu(_gn(sd, "out.0", h, 1e-5)), padding=1)


# ----------------------------------------------------------------------------- model wrapper
def apply_model(sd: SD, cfg: UNetConfig, x, sigma, context, table=None, model_dtype=None):
    """BaseModel.apply_model + EPS, LD.py:5828-5860, 1259-1265. Returns denoised (fp32).

    ``model_dtype`` mimics the reference's cast of xc/context to the UNet dtype (values are rounded
    to that dtype, arithmetic stays in the dtype of the weights).
    """
    table = make_sigmas_table() if table is None else table
    s = sigma.view(-1, 1, 1, 1)
    xc = x / (s ** 2 + 1.0) ** 0.5
    t = timestep_of_sigma(sigma, table.log_sigmas.to(sigma.device)).float()
    if model_dtype is not None:
        xc, context = xc.to(model_dtype).to(xc.dtype), context.to(model_dtype).to(context.dtype)
    eps = unet_forward(sd, cfg, xc, t, context).float()
    return x - eps * s


def cfg_combine(cond_pred, uncond_pred, scale: float):
    """cfg_function, LD.py:2594-2606 (on denoised predictions)."""
    return uncond_pred + (cond_pred - uncond_pred) * scale


def cfg_denoiser(sd: SD, cfg: UNetConfig, cond, uncond, scale: float, table=None, model_dtype=None) -> Callable:
    """model(x, sigma) -> CFG-combined denoised; one batched call ordered [uncond, cond]
    (calc_cond_batch, LD.py:2515-2569; SURVEY gotcha 7)."""
    def model(x, sigma, **_):
        b = x.shape[0]
        den = apply_model(sd, cfg, torch.cat([x, x]), torch.cat([sigma, sigma]),
                          torch.cat([uncond, cond]), table, model_dtype)
        return cfg_combine(den[b:], den[:b], scale)
    return model


# ----------------------------------------------------------------------------- samplers
def noise_scaling(sigma0, noise, latent, max_denoise: bool):
    """EPS.noise_scaling, LD.py:1267-1274."""
    noise = noise * torch.sqrt(1.0 + sigma0 ** 2.0) if max_denoise else noise * sigma0
    return noise + latent


def is_max_denoise(sigma0: float, sigma_max: float) -> bool:
    """Sampler.max_denoise, LD.py:2719-2722."""
    return math.isclose(sigma_max, sigma0, rel_tol=1e-05) or sigma0 > sigma_max


def sample_dpmpp_2m(model, x, sigmas, extra_args=None):
    """sample_dpmpp_2m_sde with eta=0, solver 'midpoint' == DPM++ 2M. LD.py:1174-1244."""
    extra_args = {} if extra_args is None else extra_args
    s_in = x.new_ones([x.shape[0]])
    old_denoised, h_last, h = None, None, None
    for i in range(len(sigmas) - 1):
        denoised = model(x, sigmas[i] * s_in, **extra_args)
        if sigmas[i + 1] == 0:
            x = denoised
        else:
            t, s = -sigmas[i].log(), -sigmas[i + 1].log()
            h = s - t
            x = sigmas[i + 1] / sigmas[i] * x + (-h).expm1().neg() * denoised
            if old_denoised is not None:
                r = h_last / h
                x = x + 0.5 * (-h).expm1().neg() * (1 / r) * (denoised - old_denoised)
        old_denoised, h_last = denoised, h
    return x


def ancestral_step(sigma_from, sigma_to, eta=1.0):
    """LD.py:844-850."""
    sigma_up = min(sigma_to, eta * (sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2) ** 0.5)
    return (sigma_to ** 2 - sigma_up ** 2) ** 0.5, sigma_up


def sample_euler_ancestral(model, x, sigmas, extra_args=None, eta=1.0, s_noise=1.0, noise_sampler=None):
    """LD.py:907-941."""
    extra_args = {} if extra_args is None else extra_args
    noise_sampler = (lambda a, b: torch.randn_like(x)) if noise_sampler is None else noise_sampler
    s_in = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        denoised = model(x, sigmas[i] * s_in, **extra_args)
        sigma_down, sigma_up = ancestral_step(sigmas[i], sigmas[i + 1], eta)
        d = (x - denoised) / sigmas[i]
        x = x + d * (sigma_down - sigmas[i])
        if sigmas[i + 1] > 0:
            x = x + noise_sampler(sigmas[i], sigmas[i + 1]) * s_noise * sigma_up
    return x


def sample_dpm_adaptive(model, x, sigma_min, sigma_max, extra_args=None, order=3, rtol=0.05, atol=0.0078, h_init=0.05,
                        pcoeff=0.0, icoeff=1.0, dcoeff=0.0, accept_safety=0.81, eta=0.0, return_info=False):
    """sample_dpm_adaptive / DPMSolver.dpm_solver_adaptive, LD.py:1049-1170, with PIDStepSizeController (LD.py:944-973):
    embedded DPM-Solver-2 (r1 = 1/3) and DPM-Solver-3 (r1 = 1/3, r2 = 2/3) sharing eps(x, s) and eps(u1, s1), an
    error norm against max(atol, rtol * max(|x_low|, |x_prev|)), and an I-controller through 1 + atan(factor - 1)."""
    extra_args = {} if extra_args is None else extra_args
    sig = lambda t: t.neg().exp()
    t_start, t_end = -torch.tensor(float(sigma_max)).log(), -torch.tensor(float(sigma_min)).log()
    forward = bool(t_end > t_start)
    h_cur = abs(h_init) * (1 if forward else -1)
    atol_t, rtol_t = torch.tensor(atol), torch.tensor(rtol)
    pid_order = 1.5 if eta else order
    b1, b2, b3 = (pcoeff + icoeff + dcoeff) / pid_order, -(pcoeff + 2 * dcoeff) / pid_order, dcoeff / pid_order
    errs = []
    s, x_prev = t_start, x
    info = {"steps": 0, "nfe": 0, "n_accept": 0, "n_reject": 0}

    def eps_at(xx, t):
        return (xx - model(xx, sig(t) * xx.new_ones([xx.shape[0]]), **extra_args)) / sig(t)

    r1, r2 = 1 / 3, 2 / 3
    while (s < t_end - 1e-5) if forward else (s > t_end + 1e-5):
        t = torch.minimum(t_end, s + h_cur) if forward else torch.maximum(t_end, s + h_cur)
        h = t - s
        eps = eps_at(x, s)
        s1, s2 = s + r1 * h, s + r2 * h
        u1 = x - sig(s1) * (r1 * h).expm1() * eps
        eps_r1 = eps_at(u1, s1)
        x_low = x - sig(t) * h.expm1() * eps - sig(t) / (2 * r1) * h.expm1() * (eps_r1 - eps)
        u2 = x - sig(s2) * (r2 * h).expm1() * eps - sig(s2) * (r2 / r1) * ((r2 * h).expm1() / (r2 * h) - 1) * (eps_r1 - eps)
        eps_r2 = eps_at(u2, s2)
        x_high = x - sig(t) * h.expm1() * eps - sig(t) / r2 * (h.expm1() / h - 1) * (eps_r2 - eps)
        delta = torch.maximum(atol_t.to(x.device), rtol_t.to(x.device) * torch.maximum(x_low.abs(), x_prev.abs()))
        error = torch.linalg.norm((x_low - x_high) / delta) / x.numel() ** 0.5
        inv_error = 1 / (float(error) + 1e-8)
        if not errs:
            errs = [inv_error, inv_error, inv_error]
        errs[0] = inv_error
        factor = 1 + math.atan(errs[0] ** b1 * errs[1] ** b2 * errs[2] ** b3 - 1)
        accept = factor >= accept_safety
        if accept:
            errs[2], errs[1] = errs[1], errs[0]
            x_prev, x, s = x_low, x_high, t
            info["n_accept"] += 1
        else:
            info["n_reject"] += 1
        h_cur *= factor
        info["nfe"] += order
        info["steps"] += 1
    return (x, info) if return_info else x


# ----------------------------------------------------------------------------- VAE decoder
def _vae_res(sd: SD, p: str, x):
    """ResnetBlock.forward, LD.py:3562-3576 (GN eps 1e-6)."""
    h = _conv(sd, p + ".conv1", F.silu(_gn(sd, p + ".norm1", x, 1e-6)), padding=1)
    h = _conv(sd, p + ".conv2", F.silu(_gn(sd, p + ".norm2", h, 1e-6)), padding=1)
    if (p + ".nin_shortcut.weight") in sd:
        x = _conv(sd, p + ".nin_shortcut", x)
    return x + h


def _vae_attn(sd: SD, p: str, x):
    """AttnBlock.forward + pytorch_attention, LD.py:3631-3642, 3591-3602 (1 head, d=C)."""
    b, c, h, w = x.shape
    hn = _gn(sd, p + ".norm", x, 1e-6)
    q, k, v = (_conv(sd, f"{p}.{n}", hn).flatten(2).transpose(1, 2) for n in ("q", "k", "v"))
    o = attention(q, k, v, 1).transpose(1, 2).reshape(b, c, h, w)
    return x + _conv(sd, p + ".proj_out", o)


def vae_decode_raw(sd: SD, cfg: VAEConfig, z):
    """AutoencodingEngine.decode + Decoder.forward, LD.py:3470-3473, 3857-3882. NCHW out."""
    h = _conv(sd, "post_quant_conv", z)
    h = _conv(sd, "decoder.conv_in", h, padding=1)
    h = _vae_res(sd, "decoder.mid.block_1", h)
    h = _vae_attn(sd, "decoder.mid.attn_1", h)
    h = _vae_res(sd, "decoder.mid.block_2", h)
    for i_level, blocks, has_up in vae_decoder_levels(cfg):
        for j in range(len(blocks)):
            h = _vae_res(sd, f"decoder.up.{i_level}.block.{j}", h)
        if has_up:
            h = _conv(sd, f"decoder.up.{i_level}.upsample.conv",
                      F.interpolate(h, scale_factor=2.0, mode="nearest"), padding=1)
    h = F.silu(_gn(sd, "decoder.norm_out", h, 1e-6))
    return _conv(sd, "decoder.conv_out", h, padding=1)


def vae_decode(sd: SD, cfg: VAEConfig, z):
    """VAE.decode, LD.py:6357-6381: clamp((x+1)/2, 0, 1), NCHW -> NHWC, fp32."""
    img = vae_decode_raw(sd, cfg, z).float()
    return torch.clamp((img + 1.0) / 2.0, min=0.0, max=1.0).movedim(1, -1)


# ----------------------------------------------------------------------------- VAE encoder
def vae_encode_moments(sd: SD, cfg: VAEConfig, x):
    """Encoder.forward + quant_conv, LD.py:3733-3758, 3478-3479. x NCHW in [-1, 1]; returns [B, 2*z, h/8, w/8].
    Downsample (LD.py:3514-3528) pads (0,1,0,1) and convolves with stride 2, padding 0."""
    h = _conv(sd, "encoder.conv_in", x, padding=1)
    for i_level, blocks, has_down in vae_encoder_levels(cfg):
        for j in range(len(blocks)):
            h = _vae_res(sd, f"encoder.down.{i_level}.block.{j}", h)
        if has_down:
            h = _conv(sd, f"encoder.down.{i_level}.downsample.conv", F.pad(h, (0, 1, 0, 1)), stride=2)
    h = _vae_res(sd, "encoder.mid.block_1", h)
    h = _vae_attn(sd, "encoder.mid.attn_1", h)
    h = _vae_res(sd, "encoder.mid.block_2", h)
    h = F.silu(_gn(sd, "encoder.norm_out", h, 1e-6))
    return _conv(sd, "quant_conv", _conv(sd, "encoder.conv_out", h, padding=1))


def posterior_sample(moments, noise):
    """DiagonalGaussianDistribution.__init__/sample, LD.py:166-179: mean + exp(0.5 * clamp(logvar, -30, 20)) * noise."""
    mean, logvar = torch.chunk(moments, 2, dim=1)
    return mean + torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0)) * noise


def vae_encode(sd: SD, cfg: VAEConfig, pixels, noise):
    """VAE.encode, LD.py:6383-6410: NHWC [0,1] pixels -> movedim -> 2x-1 -> encode -> sample -> fp32 latent.
    ``noise`` stands for the reference's torch.randn(mean.shape) drawn from the global CPU generator."""
    x = pixels.movedim(-1, 1) * 2.0 - 1.0
    return posterior_sample(vae_encode_moments(sd, cfg, x), noise).float()


# ----------------------------------------------------------------------------- latent upscale
def _slerp(b1, b2, r):
    """slerp of LD.py:430-461 on [N, C] batches with ratio r [N, 1]."""
    n1, n2 = torch.norm(b1, dim=-1, keepdim=True), torch.norm(b2, dim=-1, keepdim=True)
    u1 = torch.where(n1 == 0.0, torch.zeros_like(b1), b1 / n1)
    u2 = torch.where(n2 == 0.0, torch.zeros_like(b2), b2 / n2)
    dot = (u1 * u2).sum(1)
    omega = torch.acos(dot)
    so = torch.sin(omega)
    res = (torch.sin((1.0 - r.squeeze(1)) * omega) / so).unsqueeze(1) * u1 + (torch.sin(r.squeeze(1) * omega) / so).unsqueeze(1) * u2
    res = res * (n1 * (1.0 - r) + n2 * r)
    res = torch.where((dot > 1 - 1e-5).unsqueeze(1), b1, res)
    return torch.where((dot < 1e-5 - 1).unsqueeze(1), b1 * (1.0 - r) + b2 * r, res)


def _bilinear_table(length_old: int, length_new: int):
    """generate_bilinear_data, LD.py:464-485."""
    c1 = F.interpolate(torch.arange(length_old, dtype=torch.float32).reshape(1, 1, 1, -1), size=(1, length_new), mode="bilinear")
    ratios = (c1 - c1.floor()).reshape(-1)
    c2 = torch.arange(length_old, dtype=torch.float32).reshape(1, 1, 1, -1) + 1
    c2[:, :, :, -1] -= 1
    c2 = F.interpolate(c2, size=(1, length_new), mode="bilinear")
    return ratios, c1.to(torch.int64).reshape(-1), c2.to(torch.int64).reshape(-1)


def bislerp(samples, width: int, height: int):
    """bislerp, LD.py:429-518: slerp of the channel vectors along the width, then the height."""
    x = samples.float()
    n, c, h, w = x.shape
    r, i1, i2 = (t.to(x.device) for t in _bilinear_table(w, width))
    a = x[:, :, :, i1].movedim(1, -1).reshape(-1, c)
    b = x[:, :, :, i2].movedim(1, -1).reshape(-1, c)
    rr = r.reshape(1, 1, -1).expand(n, h, -1).reshape(-1, 1)
    x = _slerp(a, b, rr).reshape(n, h, width, c).movedim(-1, 1)
    r, i1, i2 = (t.to(x.device) for t in _bilinear_table(h, height))
    a = x[:, :, i1, :].movedim(1, -1).reshape(-1, c)
    b = x[:, :, i2, :].movedim(1, -1).reshape(-1, c)
    rr = r.reshape(1, -1, 1).expand(n, -1, width).reshape(-1, 1)
    x = _slerp(a, b, rr).reshape(n, height, width, c).movedim(-1, 1)
    return x.to(samples.dtype)


# ----------------------------------------------------------------------------- CLIP text tower
def clip_text_forward(sd: SD, cfg, tokens, intermediate_output: Optional[int] = None,
                      final_layer_norm_intermediate: bool = True):
    """CLIPTextModel.forward, LD.py:4437-4486 over CLIPEncoder / CLIPLayer / CLIPAttention / CLIPMLP (LD.py:4268-4396):
    token + position embeddings, causal mask, pre-LN layers with quick_gelu, final LayerNorm, pooled = the row at the
    arg-max token id (EOS), text_projection.  Returns (last, intermediate or None, projected pooled, pooled)."""
    tokens = tokens.long()
    x = sd["text_model.embeddings.token_embedding.weight"][tokens] + sd["text_model.embeddings.position_embedding.weight"]
    n = x.shape[1]
    mask = torch.full((n, n), float("-inf"), dtype=x.dtype, device=x.device).triu_(1)
    if intermediate_output is not None and intermediate_output < 0:
        intermediate_output = cfg.num_layers + intermediate_output
    inter = None
    for i in range(cfg.num_layers):
        p = f"text_model.encoder.layers.{i}"
        y = _ln(sd, p + ".layer_norm1", x)
        q, k, v = (_lin(sd, f"{p}.self_attn.{nm}_proj", y) for nm in ("q", "k", "v"))
        b, t, c = q.shape
        h = cfg.num_heads
        qh, kh, vh = (z.view(b, t, h, c // h).transpose(1, 2) for z in (q, k, v))
        a = F.scaled_dot_product_attention(qh, kh, vh, attn_mask=mask).transpose(1, 2).reshape(b, t, c)
        x = x + _lin(sd, p + ".self_attn.out_proj", a)
        y = _lin(sd, p + ".mlp.fc1", _ln(sd, p + ".layer_norm2", x))
        x = x + _lin(sd, p + ".mlp.fc2", y * torch.sigmoid(1.702 * y))
        if i == intermediate_output:
            inter = x.clone()
    x = _ln(sd, "text_model.final_layer_norm", x)
    if inter is not None and final_layer_norm_intermediate:
        inter = _ln(sd, "text_model.final_layer_norm", inter)
    pooled = x[torch.arange(x.shape[0], device=x.device), tokens.to(torch.int).argmax(dim=-1)]
    return x, inter, F.linear(pooled, sd["text_projection.weight"]), pooled


def clip_encode_token_weights(sd: SD, cfg, token_weight_pairs, layer_idx: Optional[int] = -2, empty_tokens=None):
    """ClipTokenWeightEncoder.encode_token_weights (LD.py:4539-4569) over SDClipModel.forward (LD.py:4689-4724) with
    layer = "hidden": one 77-token section per entry; weighted tokens are pushed away from the empty prompt's
    embedding, z = (z - z_empty) * w + z_empty; sections are concatenated along the token axis."""
    to_encode = [[t for t, _ in sec] for sec in token_weight_pairs]
    has_weights = any(w != 1.0 for sec in token_weight_pairs for _, w in sec)
    sections = len(to_encode)
    if has_weights or sections == 0:
        to_encode.append(list(empty_tokens))
    dev = sd["text_projection.weight"].device
    last, inter, pooled, _ = clip_text_forward(sd, cfg, torch.tensor(to_encode, device=dev), layer_idx)
    out = (last if layer_idx is None else inter).float()
    zs = []
    for k in range(sections):
        z = out[k:k + 1].clone()
        if has_weights:
            w = torch.tensor([wt for _, wt in token_weight_pairs[k]], device=dev, dtype=z.dtype).view(1, -1, 1)
            z = torch.where(w != 1.0, (z - out[-1]) * w + out[-1], z)
        zs.append(z)
    return torch.cat(zs, dim=-2), pooled[0:1].float()


# ----------------------------------------------------------------------------- ESRGAN upscaler
def esrgan_forward(sd: SD, x):
    """RRDBNet.forward, LD.py:7073-7140 / 7224: conv_first -> ShortcutBlock(RRDB x nb -> trunk conv) -> (nearest 2x ->
    conv -> LeakyReLU) x log2(scale) -> HR conv -> LeakyReLU -> last conv.  RRDB (LD.py:6897-6902) = three
    ResidualDenseBlock_5C (LD.py:6990-6997), each x5 * 0.2 + x over a dense 5-conv stack with LeakyReLU(0.2)
    (act(), LD.py:6767-6770); the RRDB output is again out * 0.2 + x.  x is NCHW."""
    from lightdiffusion_b200.esrgan import to_old_arch          # key renaming only (new_to_old_arch, LD.py:7142-7199)
    sd = to_old_arch(sd)
    lrelu = lambda t: F.leaky_relu(t, 0.2)
    c3 = lambda name, t: F.conv2d(t, sd[name + ".weight"], sd[name + ".bias"], padding=1)
    nb = max(int(k.split(".")[3]) for k in sd if ".RDB1.conv1.0.weight" in k) + 1
    fea = c3("model.0", x)
    h = fea
    for i in range(nb):
        r = h
        for j in (1, 2, 3):
            p = f"model.1.sub.{i}.RDB{j}"
            xs = [r]
            for k in range(1, 5):
                xs.append(lrelu(c3(f"{p}.conv{k}.0", torch.cat(xs, 1))))
            r = c3(f"{p}.conv5.0", torch.cat(xs, 1)) * 0.2 + r
        h = r * 0.2 + h
    h = fea + c3(f"model.1.sub.{nb}", h)
    tops = sorted(int(k.split(".")[1]) for k in sd if k.count(".") == 2 and k.endswith(".weight") and k != "model.0.weight")
    for n in tops[:-2]:
        h = lrelu(c3(f"model.{n}", F.interpolate(h, scale_factor=2, mode="nearest")))
    h = lrelu(c3(f"model.{tops[-2]}", h))
    return c3(f"model.{tops[-1]}", h)


# ----------------------------------------------------------------------------- TAESD
def _taesd_stack(sd: SD, layout, x):
    """Decoder2 / Encoder2 as nn.Sequential index lists (LD.py:705-723); Block = relu(conv-relu-conv-relu-conv(x) + x)."""
    c3 = lambda n, t, stride=1: F.conv2d(t, sd[n + ".weight"], sd.get(n + ".bias"), stride=stride, padding=1)
    for kind, i in layout:
        if kind == "block":
            h = c3(f"{i}.conv.4", F.relu(c3(f"{i}.conv.2", F.relu(c3(f"{i}.conv.0", x)))))
            x = F.relu(h + x)
        elif kind == "upconv":
            x = c3(str(i), F.interpolate(x, scale_factor=2, mode="nearest"))
        elif kind == "down":
            x = c3(str(i), x, 2)
        elif kind == "conv_relu":
            x = F.relu(c3(str(i), x))
        else:
            x = c3(str(i), x)
    return x


def taesd_decode(sd: SD, latent):
    """TAESD.decode, LD.py:748-753: Clamp (tanh(x/3)*3) -> Decoder2 -> .sub(0.5).mul(2); vae_shift 0, vae_scale 1."""
    from lightdiffusion_b200.taesd import DECODER_LAYOUT
    return _taesd_stack(sd, DECODER_LAYOUT, torch.tanh(latent / 3) * 3).sub(0.5).mul(2)


def taesd_encode(sd: SD, image):
    """TAESD.encode, LD.py:756-759: Encoder2(x * 0.5 + 0.5)."""
    from lightdiffusion_b200.taesd import ENCODER_LAYOUT
    return _taesd_stack(sd, ENCODER_LAYOUT, image * 0.5 + 0.5)


# ----------------------------------------------------------------------------- pipeline
LATENT_SCALE = 0.18215      # SD15 latent format, LD.py:137-139


def sample_latent(unet_sd: SD, ucfg: UNetConfig, noise, cond, uncond, cfg_scale: float, sigmas,
                  sampler: str = "dpmpp_2m", latent=None, noise_sampler=None):
    """CFGGuider.inner_sample (LD.py:2926-2965) + KSAMPLER.sample (LD.py:2738-2773):
    non-empty latent * 0.18215 on the way in, noise scaling, sampler loop, / 0.18215 on the way out."""
    table = make_sigmas_table()
    latent = torch.zeros_like(noise) if latent is None else latent
    if torch.count_nonzero(latent) > 0:
        latent = latent * LATENT_SCALE
    x = noise_scaling(sigmas[0], noise, latent, is_max_denoise(float(sigmas[0]), float(table.sigma_max)))
    model = cfg_denoiser(unet_sd, ucfg, cond, uncond, cfg_scale, table)
    if sampler == "dpmpp_2m":
        x = sample_dpmpp_2m(model, x, sigmas)
    elif sampler == "euler_ancestral":
        x = sample_euler_ancestral(model, x, sigmas, noise_sampler=noise_sampler)
    elif sampler == "dpm_adaptive":            # dpm_adaptive_function, LD.py:2779-2797
        smin = sigmas[-1] if sigmas[-1] != 0 else sigmas[-2]
        x = sample_dpm_adaptive(model, x, float(smin), float(sigmas[0]))
    else:
        raise ValueError(sampler)
    return x.float() / LATENT_SCALE


def txt2img(unet_sd: SD, ucfg: UNetConfig, vae_sd: SD, vcfg: VAEConfig, noise, cond, uncond,
            cfg_scale: float, sigmas, sampler: str = "dpmpp_2m", latent=None, noise_sampler=None):
    """common_ksampler (LD.py:6657) followed by VAEDecode (LD.py:6581). Returns (latent, image NHWC)."""
    z = sample_latent(unet_sd, ucfg, noise, cond, uncond, cfg_scale, sigmas, sampler, latent, noise_sampler)
    return z, vae_decode(vae_sd, vcfg, z)


def img2img(unet_sd: SD, ucfg: UNetConfig, enc_sd: SD, dec_sd: SD, vcfg: VAEConfig, pixels, posterior_noise, noise,
            cond, uncond, cfg_scale: float, sigmas, sampler: str = "euler_ancestral", noise_sampler=None):
    """VAEEncode (LD.py:6383) -> common_ksampler with a non-empty latent (LD.py:6657) -> VAEDecode (LD.py:6581)."""
    latent = vae_encode(enc_sd, vcfg, pixels, posterior_noise)
    return txt2img(unet_sd, ucfg, dec_sd, vcfg, noise, cond, uncond, cfg_scale, sigmas, sampler, latent, noise_sampler)


def hires_fix(unet_sd: SD, ucfg: UNetConfig, dec_sd: SD, vcfg: VAEConfig, noise, noise2, cond, uncond, cfg_scale: float,
              sigmas, sigmas2, cfg2: float = 8.0, scale: int = 2, noise_sampler=None):
    """LD.py:10048-10080: sample -> LatentUpscale(bislerp, 2x) -> euler_ancestral second pass -> VAEDecode."""
    z1 = sample_latent(unet_sd, ucfg, noise, cond, uncond, cfg_scale, sigmas, "dpmpp_2m")
    up = bislerp(z1, z1.shape[3] * scale, z1.shape[2] * scale)
    z2, img = txt2img(unet_sd, ucfg, dec_sd, vcfg, noise2, cond, uncond, cfg2, sigmas2, "euler_ancestral", up, noise_sampler)
    return z1, z2, img
