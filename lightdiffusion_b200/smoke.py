"""One small invocation of the hot path on cuda:0, checked against the oracle (used by
__graft_entry__.smoke()).  Runs every kernel family: implicit-GEMM convs, GEMMs, GroupNorm,
LayerNorm, flash attention, CFG, the DPM++ 2M update and the VAE decoder."""
import math

import torch


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
    cfg = sd15.UNetConfig(model_channels=128, channel_mult=(1, 2, 2, 2), num_heads=2, context_dim=256)
    rnd = lambda sd: {k: v.half().float().to(dev) for k, v in sd.items()}
    usd = rnd(sd15.seeded_state_dict(sd15.unet_param_shapes(cfg), 1234, bias_std=0.05, norm_jitter=0.1))
    vsd = rnd(sd15.seeded_state_dict(sd15.vae_decoder_param_shapes(), 99, bias_std=0.05, norm_jitter=0.1))
    pipe = Txt2ImgPipeline(UNetEngine(usd, cfg, dev), VAEDecoderEngine(vsd, device=dev))
    g = torch.Generator().manual_seed(0)
    b, hw = 1, 16
    noise = torch.randn(b, 4, hw, hw, generator=g).to(dev)
    cond, uncond = torch.randn(b, 77, 256, generator=g).to(dev), torch.randn(b, 77, 256, generator=g).to(dev)
    sig = calculate_sigmas("karras", 3)
    z, img = pipe.generate(noise, cond, uncond, 7.0, sig, "dpmpp_2m")
    with torch.inference_mode():
        zr, imgr = O.txt2img(usd, cfg, vsd, sd15.SD15_VAE, noise, cond.half().float(), uncond.half().float(), 7.0,
                             sig.to(dev), "dpmpp_2m")
    rel = float((z - zr).norm() / zr.norm())
    mse = float(((img - imgr) ** 2).mean())
    psnr = 10 * math.log10(1.0 / max(mse, 1e-20))
    print(f"smoke: latent rel-L2 {rel:.2e}, image PSNR {psnr:.1f} dB, launches {pipe.gpu_launches}")
    assert rel < 2e-2 and psnr >= 40.0, (rel, psnr)
