"""ORACLE — test infrastructure, not product code.

Generates tests/golden/*.pt by running the REAL reference (/root/reference, imported in the
build container through oracle/ref_loader.py) on seeded weights and inputs.  The fixtures
hold inputs and reference outputs only; weights are regenerated from the seed by
``lightdiffusion_b200.sd15.seeded_state_dict`` (CPU generator, identical on every host).

    python -m oracle.make_golden            # writes tests/golden/

Every fixture records the call that produced it (LD.py line of the entry point).
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from lightdiffusion_b200 import sd15  # noqa: E402
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

UNET_SEED, VAE_SEED = 1234, 99
INIT = dict(bias_std=0.05, norm_jitter=0.1)   # non-trivial biases / affine params


def _conds(b, ctx_dim, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(b, 77, ctx_dim, generator=g), torch.randn(b, 77, ctx_dim, generator=g)


def golden_schedules(R):
    ms = R.ModelSamplingDiscrete(R.sm_SD15({}))
    out = {
        "sigmas_table": ms.sigmas.clone(),
        "karras20": R.calculate_sigmas(ms, "karras", 20),
        "karras8": R.calculate_sigmas(ms, "karras", 8),
        "normal10": R.calculate_sigmas(ms, "normal", 10),
        "normal22": R.calculate_sigmas(ms, "normal", 22),
        "probe_sigmas": torch.tensor([14.6146, 5.0, 1.0, 0.3, 0.0292, 20.0, 0.01]),
    }
    out["probe_timesteps"] = ms.timestep(out["probe_sigmas"])
    out["probe_t"] = torch.tensor([0.0, 10.5, 500.25, 998.9, 999.0])
    out["probe_sigma_of_t"] = ms.sigma(out["probe_t"])
    out["source"] = "ModelSamplingDiscrete LD.py:1280-1351, calculate_sigmas LD.py:3045"
    torch.save(out, os.path.join(OUT, "schedules.pt"))


def golden_unet(R, tag, cfg, hw, b):
    """Reference BaseModel.apply_model (LD.py:5828) and sampling_function (LD.py:2609)."""
    sd = sd15.seeded_state_dict(sd15.unet_param_shapes(cfg), UNET_SEED, **INIT)
    model = ref_loader.build_ref_model(R, cfg, sd)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(b, 4, hw, hw, generator=g)
    ctx = torch.randn(b, 77, cfg.context_dim, generator=g)
    sig = torch.tensor([5.0, 0.3][:b] if b <= 2 else [5.0] * b)
    with torch.inference_mode():
        t = model.model_sampling.timestep(sig).float()
        xc = x / (sig.view(-1, 1, 1, 1) ** 2 + 1) ** 0.5
        eps = model.diffusion_model(xc, t, context=ctx, transformer_options={}).float()
        den = model.apply_model(x, sig, c_crossattn=ctx, transformer_options={})
        # CFG through the reference's own cond batching
        pos, neg = _conds(b, cfg.context_dim, 3)
        cfg_den = R.sampling_function(model, x, sig, R.convert_cond([[neg, {}]]), R.convert_cond([[pos, {}]]),
                                      7.0, model_options={"transformer_options": {}})
    out = dict(x=x, ctx=ctx, sigma=sig, timestep=t, eps=eps.clone(), denoised=den.clone(),
               pos=pos, neg=neg, cfg_denoised=cfg_den.clone(), cfg_scale=7.0,
               unet_seed=UNET_SEED, init=INIT, cfg=cfg,
               n_params=sum(p.numel() for p in model.diffusion_model.parameters()),
               source="UNetModel1.forward LD.py:5688, apply_model LD.py:5828, sampling_function LD.py:2609")
    torch.save(out, os.path.join(OUT, f"unet_{tag}.pt"))
    return model, sd


def golden_samplers(R, model, cfg, hw):
    """Reference sample() (LD.py:3010) with dpmpp_2m_sde eta=0 and sample1() euler_ancestral."""
    dev = torch.device("cpu")
    patcher = R.ModelPatcher(model, load_device=dev, offload_device=dev, current_device=dev)
    b = 2
    pos, neg = _conds(b, cfg.context_dim, 5)
    lat = torch.zeros(b, 4, hw, hw)
    noise = R.prepare_noise(lat, 42)
    sig = R.calculate_sigmas(model.model_sampling, "karras", 6)
    traj = []

    with torch.inference_mode():
        z = R.sample(patcher, noise, [[pos, {}]], [[neg, {}]], 7.0, dev,
                     R.ksampler("dpmpp_2m_sde", {"eta": 0.0}), sig,
                     model_options=patcher.model_options, latent_image=lat, disable_pbar=True, seed=42)
        # Euler ancestral, hires-style: normal scheduler, denoise 0.45 (LD.py:10592-10603)
        lat2 = torch.randn(b, 4, hw, hw, generator=torch.Generator().manual_seed(9)) * 0.5
        noise2 = R.prepare_noise(lat2, 43)
        ks = R.KSampler1(patcher, steps=4, device=dev, sampler="euler_ancestral", scheduler="normal",
                         denoise=0.45, model_options=patcher.model_options)
        sig2 = ks.sigmas.clone()
        torch.manual_seed(1001)       # default_noise_sampler draws torch.randn_like from the global RNG
        z2 = ks.sample(noise2, [[pos, {}]], [[neg, {}]], 8.0, latent_image=lat2, disable_pbar=True, seed=43)
    out = dict(pos=pos, neg=neg, noise=noise, sigmas=sig, z_dpmpp2m=z.clone(), cfg_scale=7.0,
               lat2=lat2, noise2=noise2, sigmas2=sig2, z_euler_a=z2.clone(), cfg_scale2=8.0,
               euler_global_seed=1001, traj=traj,
               source="sample LD.py:3010 (dpmpp_2m_sde eta=0, LD.py:1174); KSampler1.sample LD.py:3106 (euler_ancestral LD.py:907)")
    torch.save(out, os.path.join(OUT, "samplers_tiny.pt"))


def golden_sampler_toy(R):
    """The two sampler functions on a closed-form linear denoiser D(x,s) = x / (1 + s^2)."""
    def model(x, sigma, **kw):
        return x / (1.0 + sigma.view(-1, 1, 1, 1) ** 2)
    ms = R.ModelSamplingDiscrete(R.sm_SD15({}))
    sig = R.calculate_sigmas(ms, "karras", 20)
    g = torch.Generator().manual_seed(11)
    x0 = torch.randn(3, 4, 8, 8, generator=g) * sig[0]
    a = R.sample_dpmpp_2m_sde(model, x0.clone(), sig, extra_args={}, disable=True, eta=0.0,
                              noise_sampler=lambda s, sn: torch.zeros_like(x0))
    torch.manual_seed(77)
    b = R.sample_euler_ancestral(model, x0.clone(), sig, extra_args={}, disable=True)
    torch.save(dict(x0=x0, sigmas=sig, dpmpp2m=a, euler_a=b, euler_global_seed=77,
                    source="sample_dpmpp_2m_sde LD.py:1174 (eta=0), sample_euler_ancestral LD.py:907"),
               os.path.join(OUT, "samplers_toy.pt"))


def golden_vae(R):
    """Reference VAE.decode (LD.py:6357) on a small latent."""
    sd = sd15.seeded_state_dict(sd15.vae_decoder_param_shapes(), VAE_SEED, **INIT)
    vae = ref_loader.build_ref_vae(R, sd)
    z = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(5))
    with torch.inference_mode():
        img = vae.decode(z)
    n_dec = sum(p.numel() for n, p in vae.first_stage_model.named_parameters() if n.startswith("decoder."))
    torch.save(dict(z=z, image=img.clone(), vae_seed=VAE_SEED, init=INIT, n_decoder_params=n_dec,
                    source="VAE.decode LD.py:6357 -> Decoder.forward LD.py:3857"),
               os.path.join(OUT, "vae_decode.pt"))


def golden_vae_encode(R):
    """Reference VAE.encode (LD.py:6383) on small images; the posterior noise comes from the global CPU generator
    exactly as DiagonalGaussianDistribution.sample draws it (LD.py:175-179)."""
    sd = sd15.seeded_state_dict(sd15.vae_encoder_param_shapes(), VAE_SEED + 1, **INIT)
    vae = ref_loader.build_ref_vae(R, sd)
    pixels = torch.rand(2, 32, 48, 3, generator=torch.Generator().manual_seed(6))
    with torch.inference_mode():
        torch.manual_seed(2024)
        z = vae.encode(pixels)
        moments = vae.first_stage_model.quant_conv(vae.first_stage_model.encoder(pixels.movedim(-1, 1) * 2.0 - 1.0))
    torch.manual_seed(2024)
    noise = torch.randn(2, 4, 4, 6)
    n_enc = sum(p.numel() for n, p in vae.first_stage_model.named_parameters() if n.startswith("encoder."))
    torch.save(dict(pixels=pixels, z=z.clone(), moments=moments.clone(), noise=noise, global_seed=2024,
                    vae_seed=VAE_SEED + 1, init=INIT, n_encoder_params=n_enc,
                    source="VAE.encode LD.py:6383 -> Encoder.forward LD.py:3733 -> DiagonalGaussianRegularizer LD.py:3446"),
               os.path.join(OUT, "vae_encode.pt"))


def golden_bislerp(R):
    """Reference bislerp (LD.py:429) on small latents: 2x, a non-integer ratio, a downscale, and the slerp edge cases
    (zero vectors, parallel and opposite neighbours)."""
    g = torch.Generator().manual_seed(31)
    x = torch.randn(2, 4, 8, 12, generator=g)
    x[0, :, 0, 0] = 0.0                              # zero-norm vector
    x[0, :, 1, 1] = x[0, :, 1, 0] * 2.5              # parallel neighbours along w
    x[0, :, 2, 3] = -x[0, :, 2, 2]                   # opposite neighbours along w
    x[1, :, 4, 5] = x[1, :, 3, 5] * 0.5              # parallel neighbours along h
    x[1, :, 6, 7] = -1.5 * x[1, :, 5, 7]             # opposite neighbours along h
    cases = {}
    for name, (w, h) in dict(x2=(24, 16), x1p5=(18, 12), odd=(17, 11), down=(6, 4), same=(12, 8)).items():
        cases[name] = dict(width=w, height=h, out=R.bislerp(x.clone(), w, h).clone())
    x16 = torch.randn(1, 16, 5, 6, generator=g)
    cases["c16"] = dict(width=9, height=10, out=R.bislerp(x16.clone(), 9, 10).clone(), x=x16)
    torch.save(dict(x=x, cases=cases, source="bislerp LD.py:429-518"), os.path.join(OUT, "bislerp.pt"))


CLIP_SEED = 77


def _clip_tokens(g, n_words, weights=None):
    """One 77-token section in the tokenizer's layout (SDTokenizer.tokenize_with_weights, LD.py:4960-5031):
    start, words, end, pad (= end for SD1), each paired with a weight."""
    ids = torch.randint(1000, 40000, (n_words,), generator=g).tolist()
    toks = [49406] + ids + [49407] * (77 - 1 - n_words)
    w = [1.0] * 77
    for pos, val in (weights or {}).items():
        w[pos] = val
    return list(zip(toks, w))


def golden_clip(R):
    """Reference SDClipModel (LD.py:4572) with the SD1 text config and seeded weights: raw CLIPTextModel.forward
    outputs, and encode_token_weights with clip-skip -2 (CLIPSetLastLayer, LD.py:10032-10035) on weighted prompts."""
    import json
    sd = sd15.seeded_state_dict(sd15.clip_text_param_shapes(), CLIP_SEED, **INIT)
    with torch.no_grad():          # CLIPTextModel.__init__ writes text_projection in place (LD.py:4475)
        clip = R.SDClipModel(device="cpu", dtype=torch.float32, textmodel_json_config=os.path.join(
            ref_loader.REF_DIR, "_internal", "clip", "sd1_clip_config.json"))
    missing, unexpected = clip.transformer.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    g = torch.Generator().manual_seed(12)
    secs = [_clip_tokens(g, 9), _clip_tokens(g, 30, {3: 1.4, 4: 1.4, 10: 0.6}), _clip_tokens(g, 75)]
    tokens = torch.tensor([[t for t, _ in s_] for s_ in secs])
    with torch.inference_mode():
        last, inter, pooled_proj, pooled = clip.transformer(tokens, None, intermediate_output=-2,
                                                            final_layer_norm_intermediate=True)
        last_only = clip.transformer(tokens[:1], None)[0]
        clip.set_clip_options({"layer": -2})
        cond_w, pooled_w = clip.encode_token_weights([secs[1]])             # weighted: appends the empty prompt
        cond_2, pooled_2 = clip.encode_token_weights([secs[0], secs[2]])    # two unweighted sections -> 154 tokens
        clip.reset_clip_options()
        cond_last, _ = clip.encode_token_weights([secs[0]])                 # layer = "last"
    torch.save(dict(secs=secs, tokens=tokens, last=last.clone(), inter=inter.clone(), pooled_proj=pooled_proj.clone(),
                    pooled=pooled.clone(), last_only=last_only.clone(), cond_w=cond_w.clone(), pooled_w=pooled_w.clone(),
                    cond_2=cond_2.clone(), pooled_2=pooled_2.clone(), cond_last=cond_last.clone(),
                    empty=R.gen_empty_tokens(clip.special_tokens, 77), clip_seed=CLIP_SEED, init=INIT,
                    source="SDClipModel LD.py:4572, CLIPTextModel.forward LD.py:4480, encode_token_weights LD.py:4540"),
               os.path.join(OUT, "clip_text.pt"))


def golden_dpm_adaptive(R):
    """Reference sample_dpm_adaptive (LD.py:1119) on the closed-form denoiser D(x,s) = x / (1 + s^2) and on a
    non-linear one, with the accept / reject counts the PID controller produced."""
    ms = R.ModelSamplingDiscrete(R.sm_SD15({}))
    sig = R.calculate_sigmas(ms, "karras", 20)
    g = torch.Generator().manual_seed(13)
    x0 = torch.randn(3, 4, 8, 8, generator=g) * sig[0]
    out = dict(x0=x0, sigma_min=float(sig[-2]), sigma_max=float(sig[0]), cases={},
               source="sample_dpm_adaptive LD.py:1119-1170, DPMSolver LD.py:976-1116")
    models = {"linear": lambda x, sigma, **kw: x / (1.0 + sigma.view(-1, 1, 1, 1) ** 2),
              "tanh": lambda x, sigma, **kw: torch.tanh(x / (1.0 + sigma.view(-1, 1, 1, 1))) * 1.5}
    for name, model in models.items():
        for opts in (dict(), dict(rtol=0.01, atol=0.002, h_init=0.2), dict(eta=1.0, pcoeff=0.1, dcoeff=0.05)):
            xr, info = R.sample_dpm_adaptive(model, x0.clone(), out["sigma_min"], out["sigma_max"], extra_args={},
                                             disable=True, return_info=True, **opts)
            out["cases"][(name, tuple(sorted(opts.items())))] = dict(x=xr.clone(), info=dict(info))
    torch.save(out, os.path.join(OUT, "dpm_adaptive.pt"))


def golden_esrgan(R):
    """Reference RRDBNet (LD.py:7025) built from seeded old-arch state dicts (2 RRDB blocks to keep the fixture small):
    a 4x and a 2x model on small images."""
    from lightdiffusion_b200.esrgan import esrgan_param_shapes
    out = dict(cases={}, init=INIT, source="RRDBNet LD.py:7025-7225 via load_state_dict LD.py:7245")
    g = torch.Generator().manual_seed(15)
    for name, (nb, scale, seed, hw, arch) in dict(x4=(2, 4, 301, (20, 28), "realesrgan"),
                                                  x2=(3, 2, 302, (16, 16), "esrgan")).items():
        sd = sd15.seeded_state_dict(esrgan_param_shapes(nb, scale, arch=arch), seed, **INIT)
        model = R.load_state_dict(dict(sd)).eval()
        assert model.scale == scale and model.num_blocks == nb, (model.scale, model.num_blocks)
        x = torch.rand(2, 3, *hw, generator=g)
        with torch.inference_mode():
            y = model(x)
        out["cases"][name] = dict(num_blocks=nb, scale=scale, seed=seed, arch=arch, x=x, y=y.clone())
    torch.save(out, os.path.join(OUT, "esrgan.pt"))


def golden_taesd(R):
    """Reference Decoder2 / Encoder2 (LD.py:705-723) with seeded weights, wrapped the way TAESD.decode / encode do
    (LD.py:748-759; TAESD() itself insists on a weights file that is not in the checkout)."""
    from lightdiffusion_b200.taesd import taesd_param_shapes
    g = torch.Generator().manual_seed(16)
    dsd = sd15.seeded_state_dict(taesd_param_shapes("decoder"), 401, **INIT)
    esd = sd15.seeded_state_dict(taesd_param_shapes("encoder"), 402, **INIT)
    dec, enc = R.Decoder2(), R.Encoder2()
    dec.load_state_dict(dsd, strict=True)
    enc.load_state_dict(esd, strict=True)
    lat = torch.randn(2, 4, 8, 12, generator=g) * 2.5
    img = torch.rand(2, 3, 64, 48, generator=g) * 2 - 1
    with torch.inference_mode():
        y = dec(lat).sub(0.5).mul(2)
        z = enc(img * 0.5 + 0.5)
    torch.save(dict(latent=lat, decoded=y.clone(), image=img, encoded=z.clone(), seeds=(401, 402), init=INIT,
                    source="Decoder2 / Encoder2 LD.py:705-723, TAESD.decode / encode LD.py:748-759"),
               os.path.join(OUT, "taesd.pt"))


def main():
    assert ref_loader.reference_available(), "needs /root/reference (build container only)"
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    R = ref_loader.load_reference()
    if len(sys.argv) > 1:                 # e.g. `make_golden.py vae_encode`: regenerate the named fixtures only
        for name in sys.argv[1:]:
            globals()["golden_" + name](R)
        return
    golden_schedules(R)
    golden_sampler_toy(R)
    tiny = sd15.tiny_unet_config()
    model, _ = golden_unet(R, "tiny", tiny, hw=16, b=2)
    golden_samplers(R, model, tiny, hw=16)
    del model
    golden_vae(R)
    golden_vae_encode(R)
    golden_bislerp(R)
    golden_clip(R)
    golden_dpm_adaptive(R)
    golden_esrgan(R)
    golden_taesd(R)
    golden_unet(R, "sd15_lat8", sd15.SD15_UNET, hw=8, b=2)   # real SD1.5 dims, small latent
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
