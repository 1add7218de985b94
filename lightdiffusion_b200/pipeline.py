"""txt2img hot path: noise -> 20-step sampler (CFG UNet step + fused update) -> VAE decode, and its
data-parallel launcher (independent images sharded over ranks; one final gather of the images).

Mirrors common_ksampler (LD.py:6657) -> CFGGuider.inner_sample (LD.py:2926-2965) -> KSAMPLER.sample
(LD.py:2738-2773) -> sampler loop, then VAEDecode (LD.py:6581).  Everything on the device is a
libldx.so kernel; the host only sequences launches.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _lib as L
from .engine import UNetEngine, VAEDecoderEngine, VAEEncoderEngine, sigma_tables
from .upscale import bislerp
from .samplers import dpmpp2m_coefficients, get_ancestral_step, get_sigmas_karras, sample_dpm_adaptive

LATENT_SCALE = 0.18215          # SD15 latent format (LD.py:137-139)


def calculate_sigmas(scheduler: str, steps: int) -> torch.Tensor:
    """calculate_sigmas (LD.py:3045-3054): 'karras' or 'normal' over the SD1.x discrete schedule."""
    sig, log_sig = sigma_tables()
    if scheduler == "karras":
        return get_sigmas_karras(steps, float(sig[0]), float(sig[-1]))
    if scheduler == "normal":           # normal_scheduler LD.py:2639-2651 with timestep()/sigma() of LD.py:1336-1351
        def timestep(s):
            return (s.log() - log_sig[:, None]).abs().argmin(dim=0).view(s.shape)
        ts = torch.linspace(timestep(sig[-1]), timestep(sig[0]), steps)
        out = []
        for t in ts:
            t = torch.clamp(t.float(), min=0, max=len(sig) - 1)
            lo, hi, w = t.floor().long(), t.ceil().long(), t.frac()
            out.append(float(((1 - w) * log_sig[lo] + w * log_sig[hi]).exp()))
        return torch.FloatTensor(out + [0.0])
    raise ValueError(f"unknown scheduler {scheduler!r}")


def sigmas_for_denoise(scheduler: str, steps: int, denoise: Optional[float]) -> torch.Tensor:
    """KSampler1.set_steps (LD.py:3097-3104)."""
    if denoise is None or denoise > 0.9999:
        return calculate_sigmas(scheduler, steps)
    return calculate_sigmas(scheduler, int(steps / denoise))[-(steps + 1):]


class Txt2ImgPipeline:
    """Resident UNet + VAE decoder on one GPU.  ``sample`` returns the latent (already / 0.18215, like
    the reference's sample()); ``generate`` also decodes to [b, 8h, 8w, 3] fp32 images in [0, 1]."""

    def __init__(self, unet: UNetEngine, vae: Optional[VAEDecoderEngine], vae_encoder: Optional[VAEEncoderEngine] = None):
        self.unet, self.vae, self.vae_encoder = unet, vae, vae_encoder
        self.sigma_max = float(unet.sigmas[-1])
        self.gpu_launches = 0            # kernels of ours launched by the last call

    @torch.no_grad()
    def sample(self, noise: torch.Tensor, cond: torch.Tensor, uncond: torch.Tensor, cfg_scale: float,
               sigmas: torch.Tensor, sampler: str = "dpmpp_2m", latent: Optional[torch.Tensor] = None,
               noise_sampler=None) -> torch.Tensor:
        b, _, H, W = noise.shape
        dev = self.unet.dev
        ctx = self.unet.cat_contexts(uncond.to(dev), cond.to(dev))     # [uncond, cond] order, LD.py:2515; lcm-padded
        plan = self.unet.plan(2 * b, H, W, cfg_pairs=True, ctx_len=ctx.shape[1])
        sig = torch.as_tensor(sigmas).detach().to("cpu", torch.float32)
        lib, st = L.lib(), torch.cuda.current_stream().cuda_stream
        launches = 0
        # noise scaling (EPS.noise_scaling LD.py:1267-1274; max_denoise LD.py:2719-2722)
        x = plan.x_in
        noise = noise.to(dev, torch.float32)
        max_denoise = math.isclose(self.sigma_max, float(sig[0]), rel_tol=1e-05) or float(sig[0]) > self.sigma_max
        x.copy_(noise * torch.sqrt(1.0 + sig[0] ** 2.0) if max_denoise else noise * sig[0])
        if latent is not None and torch.count_nonzero(latent) > 0:
            x.add_(latent.to(dev, torch.float32) * LATENT_SCALE)      # process_latent_in (LD.py:2938-2941)
        plan.set_context(ctx)
        launches += 1
        n = x.numel()
        if sampler == "dpm_adaptive":
            # dpm_adaptive_function (LD.py:2779-2797): sigma_min is the last non-zero sigma, sigma_max the first
            if len(sig) <= 1:
                return x / LATENT_SCALE
            smin = sig[-1] if sig[-1] != 0 else sig[-2]
            calls = [0]

            def guided(xx, sigma_vec, **_):
                calls[0] += 1
                return plan.run(xx, sigma_vec, None, cfg_scale=cfg_scale)

            xs, info = sample_dpm_adaptive(guided, x, float(smin), float(sig[0]), return_info=True)
            self.gpu_launches = launches + calls[0] * plan.num_launches + info["steps"] * 9
            self.sampler_info = info
            return xs / LATENT_SCALE
        old = torch.empty_like(x)
        have_old, h_last = False, None
        C = self.unet.cfg.out_channels
        if sampler == "dpmpp_2m":
            # per step: one graph replay of the UNet program + ONE fused tail launch (CFG, denoised, DPM++ 2M update,
            # old <- denoised, next sigma); coefficients in the reference's 0-d fp32 tensor arithmetic on the host
            if len(sig) > 1:
                plan.sigma_in.fill_(float(sig[0]))
            for i in range(len(sig) - 1):
                plan.run_body()
                last = bool(sig[i + 1] == 0)
                if last:
                    c_x, c_d, c_o, h = 0.0, 1.0, None, None
                else:
                    c_x, c_d, c_o, h = dpmpp2m_coefficients(sig, i, h_last if have_old else None)
                L.check(lib.ldx_cfg_dpmpp2m_step(plan.eps_buf.data_ptr(), plan.eps_buf.shape[-1], x.data_ptr(),
                                                 old.data_ptr(), None, float(sig[i]), float(cfg_scale), c_x, c_d,
                                                 0.0 if c_o is None else c_o, int(have_old and c_o is not None),
                                                 int(last), float(sig[i + 1]), plan.sigma_in.data_ptr(), b, C, H, W,
                                                 st), "ldx_cfg_dpmpp2m_step")
                launches += plan.prog.num_launches + 1
                have_old, h_last = True, h
            self.gpu_launches = launches
            return x / LATENT_SCALE
        for i in range(len(sig) - 1):
            plan.sigma_in.fill_(float(sig[i]))
            den = plan.run(x, plan.sigma_in, None, cfg_scale=cfg_scale)
            launches += plan.num_launches
            if sampler == "euler_ancestral":
                sigma_down, sigma_up = get_ancestral_step(sig[i], sig[i + 1])
                nz = None
                if sig[i + 1] > 0:
                    nz = (torch.randn_like(x) if noise_sampler is None else
                          noise_sampler(sig[i], sig[i + 1]).to(dev, torch.float32).contiguous())
                L.check(lib.ldx_euler_ancestral_step(x.data_ptr(), den.data_ptr(), None if nz is None else nz.data_ptr(),
                                                     float(sig[i]), float(sigma_down), float(sigma_up), n, st),
                        "ldx_euler_ancestral_step")
                launches += 1
            else:
                raise ValueError(f"unknown sampler {sampler!r}")
        self.gpu_launches = launches
        return x / LATENT_SCALE                                      # process_latent_out (LD.py:2965)

    @torch.no_grad()
    def generate(self, noise, cond, uncond, cfg_scale, sigmas, sampler="dpmpp_2m", latent=None, noise_sampler=None,
                 u8: bool = False):
        """``u8``: also quantise the decoded images to uint8 (what the reference saves, LD.py:6750) and return
        them instead of the fp32 images -- the compact form the multi-GPU gather moves."""
        z = self.sample(noise, cond, uncond, cfg_scale, sigmas, sampler, latent, noise_sampler)
        img = self.vae.decode_u8(z) if u8 else self.vae.decode(z)
        self.gpu_launches += self.vae.plan(*[z.shape[0], z.shape[2], z.shape[3]]).prog.num_launches + int(u8)
        return z, img


    @torch.no_grad()
    def img2img(self, pixels, noise, cond, uncond, cfg_scale, steps: int, denoise: float, scheduler="normal",
                sampler="euler_ancestral", posterior_noise=None, noise_sampler=None):
        """VAEEncode (VAE.encode, LD.py:6383-6410) -> common_ksampler with denoise < 1 (KSampler1.set_steps,
        LD.py:3097-3104: the last steps+1 sigmas of int(steps/denoise)) -> VAEDecode: the img2img / hires-fix second
        pass.  pixels fp32 [b, H, W, 3] in [0,1]; returns (latent, image)."""
        if self.vae_encoder is None:
            raise L.LdxError("img2img needs a VAEEncoderEngine")
        latent = self.vae_encoder.encode(pixels.to(self.unet.dev, torch.float32), noise=posterior_noise)
        enc_launches = self.vae_encoder.plan(*pixels.shape[:3]).prog.num_launches + 1
        sig = sigmas_for_denoise(scheduler, steps, denoise)
        z, img = self.generate(noise, cond, uncond, cfg_scale, sig, sampler, latent, noise_sampler)
        self.gpu_launches += enc_launches
        return z, img


    @torch.no_grad()
    def hires_fix(self, noise, noise2, cond, uncond, cfg_scale, sigmas, sampler="dpmpp_2m", scale: int = 2,
                  steps2: int = 10, cfg2: float = 8.0, denoise2: float = 0.45, scheduler2="normal",
                  sampler2="euler_ancestral", noise_sampler=None):
        """The reference's two-pass flow (LD.py:10048-10080): base-resolution sample -> LatentUpscale with bislerp to
        ``scale`` x (LD.py:6639-6654) -> second KSampler pass (10 Euler-ancestral steps, 'normal' schedule, denoise
        0.45, cfg 8) seeded by ``noise2`` -> VAE decode.  Returns (base latent, final latent, image)."""
        z1 = self.sample(noise, cond, uncond, cfg_scale, sigmas, sampler)
        launches = self.gpu_launches
        up = bislerp(z1, z1.shape[3] * scale, z1.shape[2] * scale)
        launches += 2
        z2, img = self.generate(noise2, cond, uncond, cfg2, sigmas_for_denoise(scheduler2, steps2, denoise2), sampler2,
                                up, noise_sampler)
        self.gpu_launches += launches
        return z1, z2, img


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous block of images for this rank (independent units; no collective on the step path)."""
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)


def gather_images(local: torch.Tensor, n_total: int, group=None) -> Optional[torch.Tensor]:
    """The one collective of the path: every rank's decoded images land on rank 0 ONLY (a gather over
    NCCL/NVLink on GPUs, gloo in CPU tests); the other ranks send their block and move on.  ``local`` is this
    rank's [n_local, H, W, 3] block (uint8 from ``decode_u8`` -- 786 KB per 512x512 image -- or fp32), padded to
    the common shard size.  Returns the [n_total, H, W, 3] stack on rank 0 and None elsewhere."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = (n_total + world - 1) // world
    pad = per - local.shape[0]
    if pad:
        local = torch.cat([local, local.new_zeros(pad, *local.shape[1:])])
    local = local.contiguous()
    out = [torch.empty_like(local) for _ in range(world)] if rank == 0 else None
    dist.gather(local, out, dst=0, group=group)
    if rank != 0:
        return None
    return torch.cat(out)[:n_total]
