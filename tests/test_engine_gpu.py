"""Model-level parity of the CUDA engines against the oracle restatement (itself pinned to the real
reference by tests/test_oracle.py) on the same seeded weights, fp16-rounded as the reference's fp16
model stores them.  Tolerances are north_star's: per-step UNet eps within 2e-3 relative L2, decoded
image >= 40 dB PSNR."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from lightdiffusion_b200 import sd15
    from lightdiffusion_b200.engine import UNetEngine, VAEDecoderEngine, VAEEncoderEngine
    from oracle import ldm_oracle as O
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

DEV = "cuda"
EPS_TOL = 2e-3          # north_star: per-step UNet epsilon, relative L2, fp16
PSNR_MIN = 40.0         # north_star: final decoded image


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm())


def psnr(a, b):
    mse = float(((a.double() - b.double()) ** 2).mean())
    return 10 * math.log10(1.0 / max(mse, 1e-20))


def rounded(sd):
    """fp16-round the weights once; oracle (fp32 math) and engine then share identical parameters."""
    return {k: v.to(torch.float16).float().to(DEV) for k, v in sd.items()}


SMALL = None


def small_cfg():
    # same topology as SD1.5, head dims 64 / 128 / 128 (supported by the attention kernel)
    return sd15.UNetConfig(model_channels=128, channel_mult=(1, 2, 2, 2), num_heads=2, context_dim=256)


@pytest.fixture(scope="module")
def small_unet():
    cfg = small_cfg()
    sd = rounded(sd15.seeded_state_dict(sd15.unet_param_shapes(cfg), 1234, bias_std=0.05, norm_jitter=0.1))
    return cfg, sd, UNetEngine(sd, cfg)


@pytest.fixture(scope="module")
def sd15_unet():
    cfg = sd15.SD15_UNET
    sd = rounded(sd15.seeded_state_dict(sd15.unet_param_shapes(cfg), 1234, bias_std=0.05, norm_jitter=0.1))
    return cfg, sd, UNetEngine(sd, cfg)


def _inputs(b, hw, ctx_dim, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(b, 4, hw, hw, generator=g).to(DEV)
    ctx = torch.randn(b, 77, ctx_dim, generator=g).to(DEV)
    return x, ctx


def _oracle_eps(sd, cfg, x, sigma, ctx):
    table = O.make_sigmas_table()
    xc = (x / (sigma.view(-1, 1, 1, 1) ** 2 + 1) ** 0.5).to(torch.float16).float()   # reference casts to fp16
    t = O.timestep_of_sigma(sigma.cpu(), table.log_sigmas).float().to(DEV)
    with torch.inference_mode():
        return O.unet_forward(sd, cfg, xc, t, ctx.to(torch.float16).float())


@pytest.mark.parametrize("hw,b", [(32, 2), (16, 3), (8, 1)])
def test_unet_small_eps_parity(small_unet, hw, b):
    cfg, sd, eng = small_unet
    x, ctx = _inputs(b, hw, cfg.context_dim, 7)
    for s in (14.6146, 1.99, 0.0485):
        sigma = torch.full((b,), s, device=DEV)
        xs = x * math.sqrt(1 + s * s)
        e = eng.eps(xs, sigma, ctx)
        ref = _oracle_eps(sd, cfg, xs, sigma, ctx)
        assert rel(e, ref) < EPS_TOL, (hw, b, s, rel(e, ref))


@pytest.mark.parametrize("h,w,b", [(24, 40, 2), (8, 56, 1), (72, 88, 1), (96, 64, 1)])
def test_unet_small_eps_parity_nonsquare(small_unet, h, w, b):
    """Portrait / landscape latents (512x768 is 64x96): widths that are not powers of two, token counts that are
    not multiples of the 128-query / 64-key attention tiles (24x40 -> 960, 240, 60, 15 tokens per level)."""
    cfg, sd, eng = small_unet
    g = torch.Generator().manual_seed(17)
    x = torch.randn(b, 4, h, w, generator=g).to(DEV)
    ctx = torch.randn(b, 77, cfg.context_dim, generator=g).to(DEV)
    for s in (7.5, 0.3):
        sigma = torch.full((b,), s, device=DEV)
        xs = x * math.sqrt(1 + s * s)
        e = eng.eps(xs, sigma, ctx)
        ref = _oracle_eps(sd, cfg, xs, sigma, ctx)
        assert e.shape == (b, 4, h, w)
        assert rel(e, ref) < EPS_TOL, (h, w, b, s, rel(e, ref))
    with pytest.raises(Exception):
        eng.eps(torch.zeros(1, 4, 20, 24, device=DEV), torch.ones(1, device=DEV), ctx[:1])   # 20 is not a multiple of 8


def test_unet_small_graph_equals_eager_and_is_deterministic(small_unet):
    cfg, sd, eng = small_unet
    x, ctx = _inputs(2, 16, cfg.context_dim, 8)
    sigma = torch.tensor([3.0, 0.5], device=DEV)
    pl = eng.plan(2, 16, 16)
    pl.use_graph = False
    a = pl.run(x, sigma, ctx).clone()
    pl.use_graph = True
    b1 = pl.run(x, sigma, ctx).clone()
    b2 = pl.run(x, sigma, None).clone()         # context reused
    assert torch.equal(a, b1) and torch.equal(b1, b2)


def test_unet_small_shard_invariance(small_unet):
    """Images are independent: a 4-image batch and two 2-image batches give bit-identical results
    (SURVEY.md §4 multi-GPU-without-a-cluster test; what sharding across ranks relies on)."""
    cfg, sd, eng = small_unet
    x, ctx = _inputs(4, 16, cfg.context_dim, 9)
    sigma = torch.tensor([5.0, 5.0, 5.0, 5.0], device=DEV)
    full = eng.denoise(x, sigma, ctx)
    lo = eng.denoise(x[:2], sigma[:2], ctx[:2])
    hi = eng.denoise(x[2:], sigma[2:], ctx[2:])
    assert torch.equal(full, torch.cat([lo, hi]))


def test_unet_small_cfg_matches_oracle(small_unet):
    cfg, sd, eng = small_unet
    x, _ = _inputs(2, 16, cfg.context_dim, 10)
    g = torch.Generator().manual_seed(3)
    cond, uncond = torch.randn(2, 77, cfg.context_dim, generator=g).to(DEV), torch.randn(2, 77, cfg.context_dim, generator=g).to(DEV)
    sigma = torch.tensor([7.0, 0.9], device=DEV)
    xs = x * 5
    out = eng.denoise_cfg(xs, sigma, cond, uncond, 7.0)
    sd_cpu = sd
    model = O.cfg_denoiser(sd_cpu, cfg, cond, uncond, 7.0, model_dtype=torch.float16)
    with torch.inference_mode():
        # oracle in fp32 on the GPU; sigma table lookups happen on CPU inside apply_model
        ref = model(xs, sigma)
    # CFG amplifies the cond/uncond difference by 7, so the bound is on the eps scale
    assert rel((xs - out) / sigma.view(-1, 1, 1, 1), (xs - ref) / sigma.view(-1, 1, 1, 1)) < 3 * EPS_TOL


@pytest.mark.parametrize("nb", [1, 3])
def test_cfg_shared_prefix_is_bit_identical(small_unet, nb, monkeypatch):
    """Evaluating the layers before the first cross-attention once per latent and duplicating them to the
    [uncond, cond] batch must give exactly the result of evaluating all 2b samples (calc_cond_batch feeds both
    halves the same x and timestep, LD.py:2545-2551)."""
    from lightdiffusion_b200 import engine as E
    cfg, sd, eng = small_unet
    g = torch.Generator().manual_seed(31 + nb)
    x = (torch.randn(nb, 4, 16, 16, generator=g) * 4).to(DEV)
    ctx = torch.randn(2 * nb, 77, cfg.context_dim, generator=g).to(DEV)
    sigma = torch.linspace(0.5, 6.0, nb).to(DEV)
    outs = []
    for share in (True, False):
        monkeypatch.setattr(E, "SHARE_CFG_PREFIX", share)
        plan = E.UNetPlan(eng, 2 * nb, 16, 16, True)          # fresh plan, not the engine's cache
        plan.use_graph = False
        plan.set_context(ctx)
        outs.append(plan.run(x, sigma, None, cfg_scale=7.0).clone())
    assert torch.equal(outs[0], outs[1])


def test_unet_sd15_eps_parity(sd15_unet):
    cfg, sd, eng = sd15_unet
    x, ctx = _inputs(2, 64, cfg.context_dim, 11)
    errs = []
    for s in (14.6146, 4.4998, 0.7909, 0.0292):
        sigma = torch.full((2,), s, device=DEV)
        xs = x * math.sqrt(1 + s * s)
        e = eng.eps(xs, sigma, ctx)
        ref = _oracle_eps(sd, cfg, xs, sigma, ctx)
        errs.append(rel(e, ref))
    print("SD1.5 eps rel-L2 per sigma:", errs)
    assert max(errs) < EPS_TOL, errs


def test_unet_sd15_eps_parity_benched_config_b8(sd15_unet):
    """BASELINE.json config 2, the shape bench.py times: 8 images -> the B = 16 classifier-free-guidance plan
    (different tile / CTA-pair choices than B = 2, split-K on the 8x8 convs, the shared CFG prefix at 320-channel
    width).  Every one of the 16 eps maps is compared with the oracle at four sigmas of the Karras schedule."""
    cfg, sd, eng = sd15_unet
    b = 8
    g = torch.Generator().manual_seed(23)
    x = torch.randn(b, 4, 64, 64, generator=g).to(DEV)
    cond = torch.randn(b, 77, cfg.context_dim, generator=g).to(DEV)
    uncond = torch.randn(b, 77, cfg.context_dim, generator=g).to(DEV)
    plan = eng.plan(2 * b, 64, 64, cfg_pairs=True)
    ctx = torch.cat([uncond, cond])
    plan.set_context(ctx)
    worst = []
    for s in (14.6146, 4.4998, 0.7909, 0.0292):
        sigma = torch.full((b,), s, device=DEV)
        xs = x * math.sqrt(1 + s * s)
        den = plan.run(xs, sigma, None, cfg_scale=7.0).clone()
        e = plan.eps_buf[..., :cfg.out_channels].float().permute(0, 3, 1, 2).contiguous()      # [16, 4, 64, 64]
        refs = []
        for lo in range(0, 2 * b, 4):                   # oracle in chunks of 4 samples (fp32 attention memory)
            xin = torch.cat([xs, xs])[lo:lo + 4]
            refs.append(_oracle_eps(sd, cfg, xin, torch.full((4,), s, device=DEV), ctx[lo:lo + 4]))
        ref = torch.cat(refs)
        per = [rel(e[i], ref[i]) for i in range(2 * b)]
        worst.append(max(per))
        # the fused CFG tail on top of those eps maps (cfg_function LD.py:2605 on the denoised predictions)
        du, dc = xs - ref[:b] * s, xs - ref[b:] * s
        den_ref = du + (dc - du) * 7.0
        assert rel((xs - den) / s, (xs - den_ref) / s) < 3 * EPS_TOL
    print("SD1.5 B=16 CFG plan, worst per-sample eps rel-L2 per sigma:", worst)
    assert max(worst) < EPS_TOL, worst


def test_unet_long_context_154_tokens(small_unet):
    """Two 77-token prompt sections: the cond is [b, 154, C] and the uncond [b, 77, C]; the reference repeats the
    shorter one to the lcm (CONDCrossAttn.concat, LD.py:647-663), which leaves softmax attention unchanged."""
    cfg, sd, eng = small_unet
    g = torch.Generator().manual_seed(61)
    x = (torch.randn(2, 4, 16, 16, generator=g) * 4).to(DEV)
    cond = torch.randn(2, 154, cfg.context_dim, generator=g).to(DEV)
    uncond = torch.randn(2, 77, cfg.context_dim, generator=g).to(DEV)
    sigma = torch.tensor([3.0, 0.7], device=DEV)
    out = eng.denoise_cfg(x, sigma, cond, uncond, 7.0)
    model = O.cfg_denoiser(sd, cfg, cond, uncond.repeat(1, 2, 1), 7.0, model_dtype=torch.float16)
    with torch.inference_mode():
        ref = model(x, sigma)
    sv = sigma.view(-1, 1, 1, 1)
    assert rel((x - out) / sv, (x - ref) / sv) < 3 * EPS_TOL
    e = eng.eps(x, sigma, cond)                          # plain (non-CFG) plan with a 154-token context
    assert rel(e, _oracle_eps(sd, cfg, x, sigma, cond)) < EPS_TOL
    with pytest.raises(Exception):
        eng.plan(2, 16, 16).set_context(cond)            # a 77-token plan rejects a 154-token context loudly


def test_plan_cache_is_bounded(small_unet):
    """An img2img flow over many tile shapes must not keep every plan (pool + CUDA graph) alive."""
    from lightdiffusion_b200 import engine as E
    cfg, sd, eng = small_unet
    for hw in (8, 16, 24, 32, 40, 48):
        eng.plan(1, hw, hw)
    assert len(eng._plans) <= E.MAX_PLANS
    assert (1, 48, 48, False, 77) in eng._plans and (1, 8, 8, False, 77) not in eng._plans


def test_unet_sd15_eps_parity_hires_128(sd15_unet):
    """BASELINE.json config 4: the hires-fix second pass runs the UNet on a 128x128 latent
    (16384-token self-attention at the 320-channel level)."""
    cfg, sd, eng = sd15_unet
    x, ctx = _inputs(1, 128, cfg.context_dim, 21)
    s = 2.0
    sigma = torch.full((1,), s, device=DEV)
    xs = x * math.sqrt(1 + s * s)
    e = eng.eps(xs, sigma, ctx)
    ref = _oracle_eps(sd, cfg, xs, sigma, ctx)
    print("SD1.5 128x128 eps rel-L2:", rel(e, ref))
    assert rel(e, ref) < EPS_TOL


def test_unet_sd15_denoised(sd15_unet):
    cfg, sd, eng = sd15_unet
    x, ctx = _inputs(1, 32, cfg.context_dim, 12)
    sigma = torch.tensor([2.6408], device=DEV)
    xs = x * 2.8
    den = eng.denoise(xs, sigma, ctx)
    with torch.inference_mode():
        ref = O.apply_model(sd, cfg, xs, sigma, ctx, model_dtype=torch.float16)
    assert rel((xs - den) / sigma, (xs - ref) / sigma) < EPS_TOL


@pytest.fixture(scope="module")
def vae():
    sd = rounded(sd15.seeded_state_dict(sd15.vae_decoder_param_shapes(), 99, bias_std=0.05, norm_jitter=0.1))
    return sd, VAEDecoderEngine(sd)


# (1, 128) is BASELINE.json config 5: a 1024x1024 decode
@pytest.mark.parametrize("b,hw", [(1, 16), (2, 12), (2, 32), (1, 64), (1, 128)])
def test_vae_decode_parity(vae, b, hw):
    sd, eng = vae
    z = torch.randn(b, 4, hw, hw, generator=torch.Generator().manual_seed(5)).to(DEV)
    img = eng.decode(z)
    raw = eng.decode_raw(z)
    with torch.inference_mode():
        ref_raw = O.vae_decode_raw(sd, sd15.SD15_VAE, z.half().float())
        ref = torch.clamp((ref_raw + 1.0) / 2.0, 0.0, 1.0).movedim(1, -1)
    assert img.shape == (b, hw * 8, hw * 8, 3)
    print("vae raw rel", rel(raw, ref_raw), "psnr", psnr(img, ref))
    assert rel(raw, ref_raw) < 5e-3
    assert psnr(img, ref) >= PSNR_MIN


# ------------------------------------------------------------------------------------ VAE encoder
LATENT_TOL = 5e-3       # encoded latent / posterior moments, relative L2, fp16 activations vs the fp32 oracle


@pytest.fixture(scope="module")
def vae_enc():
    sd = rounded(sd15.seeded_state_dict(sd15.vae_encoder_param_shapes(), 100, bias_std=0.05, norm_jitter=0.1))
    return sd, VAEEncoderEngine(sd)


def test_vae_encode_golden(golden):
    """Against the reference's own VAE.encode output (tests/golden/vae_encode.pt, same seeded weights)."""
    g = golden("vae_encode.pt")
    sd = sd15.seeded_state_dict(sd15.vae_encoder_param_shapes(), g["vae_seed"], **g["init"])
    eng = VAEEncoderEngine({k: v.to(DEV) for k, v in sd.items()})
    mom = eng.moments(g["pixels"].to(DEV))
    z = eng.encode(g["pixels"].to(DEV), noise=g["noise"].to(DEV))
    print("vae encode golden: moments rel", rel(mom.cpu(), g["moments"]), "z rel", rel(z.cpu(), g["z"]))
    assert z.shape == g["z"].shape and z.dtype == torch.float32
    assert rel(mom.cpu(), g["moments"]) < LATENT_TOL
    assert rel(z.cpu(), g["z"]) < LATENT_TOL
    # default noise: the same global-generator draw the reference makes
    torch.manual_seed(g["global_seed"])
    z2 = eng.encode(g["pixels"].to(DEV))
    assert torch.equal(z2, z)
    assert rel(eng.encode(g["pixels"].to(DEV), sample=False).cpu(), g["moments"][:, :4]) < LATENT_TOL


# (1, 1024, 1024) is BASELINE.json config 5 (img2img at 1024x1024)
@pytest.mark.parametrize("b,h,w", [(1, 64, 64), (2, 128, 192), (1, 512, 512), (1, 1024, 1024)])
def test_vae_encode_parity(vae_enc, b, h, w):
    sd, eng = vae_enc
    g = torch.Generator().manual_seed(7)
    # smooth image + noise, in [0,1]
    pix = (torch.rand(b, h // 8, w // 8, 3, generator=g).repeat_interleave(8, 1).repeat_interleave(8, 2) * 0.8
           + 0.2 * torch.rand(b, h, w, 3, generator=g)).to(DEV)
    noise = torch.randn(b, 4, h // 8, w // 8, generator=g).to(DEV)
    z = eng.encode(pix, noise=noise)
    mom = eng.moments(pix)
    with torch.inference_mode():
        x16 = (pix.movedim(-1, 1) * 2.0 - 1.0).half().float()
        ref_mom = O.vae_encode_moments(sd, sd15.SD15_VAE, x16)
        ref = O.posterior_sample(ref_mom, noise)
    print("vae encode rel: moments", rel(mom, ref_mom), "z", rel(z, ref))
    assert z.shape == (b, 4, h // 8, w // 8)
    assert rel(mom, ref_mom) < LATENT_TOL
    assert rel(z, ref) < LATENT_TOL


def test_vae_encode_rejects_bad_sizes(vae_enc):
    _, eng = vae_enc
    with pytest.raises(Exception):
        eng.encode(torch.zeros(1, 60, 64, 3, device=DEV))
    with pytest.raises(Exception):
        eng.encode(torch.zeros(64, 64, 3, device=DEV))


# ------------------------------------------------------------------------------------ CLIP text tower
COND_TOL = 5e-3         # conditioning (LayerNorm'd hidden states) relative L2, fp16 activations vs fp32 reference


@pytest.fixture(scope="module")
def clip_eng(golden):
    from lightdiffusion_b200.clip import CLIPTextEngine
    g = golden("clip_text.pt")
    sd = rounded(sd15.seeded_state_dict(sd15.clip_text_param_shapes(), g["clip_seed"], **g["init"]))
    return g, sd, CLIPTextEngine(sd)


def test_clip_text_forward_matches_reference_golden(clip_eng):
    """CLIPTextModel.forward outputs of the real reference (unrounded fp32 weights) vs the engine (fp16 weights)."""
    g, sd, eng = clip_eng
    last, inter, proj, pooled = eng.forward(g["tokens"], intermediate_output=-2)
    errs = dict(last=rel(last.cpu(), g["last"]), inter=rel(inter.cpu(), g["inter"]), proj=rel(proj.cpu(), g["pooled_proj"]),
                pooled=rel(pooled.cpu(), g["pooled"]))
    print("clip vs reference golden", errs)
    assert max(errs.values()) < COND_TOL
    assert last.dtype == torch.float32 and last.shape == (3, 77, 768)
    assert eng.forward(g["tokens"][:1])[1] is None
    assert rel(eng.forward(g["tokens"][:1])[0].cpu(), g["last_only"]) < COND_TOL


def test_clip_text_parity_with_oracle(clip_eng):
    """Same fp16-rounded weights on both sides; other batch sizes and intermediate layers."""
    g, sd, eng = clip_eng
    gen = torch.Generator().manual_seed(3)
    for B, T, inter in ((1, 77, -2), (4, 77, None), (2, 77, 5), (7, 77, -1)):
        tok = torch.randint(0, 49408, (B, T), generator=gen)
        tok[:, 0] = 49406
        last, mid, proj, pooled = eng.forward(tok, intermediate_output=inter)
        with torch.inference_mode():
            rl, rm, rp, rpo = O.clip_text_forward(sd, sd15.SD15_CLIP, tok.to(DEV), inter)
        e = [rel(last, rl), rel(proj, rp), rel(pooled, rpo)] + ([] if inter is None else [rel(mid, rm)])
        print("clip parity", (B, T, inter), e)
        assert max(e) < 3e-3
    raw = eng.forward(g["tokens"], intermediate_output=-2, final_layer_norm_intermediate=False)[1]
    with torch.inference_mode():
        rr = O.clip_text_forward(sd, sd15.SD15_CLIP, g["tokens"].to(DEV), -2, False)[1]
    assert rel(raw, rr) < 3e-3
    with pytest.raises(Exception):
        eng.forward(torch.full((1, 77), 49408))          # id outside the vocabulary
    with pytest.raises(Exception):
        eng.forward(torch.zeros(1, 78, dtype=torch.long))
    with pytest.raises(Exception):
        eng.forward(torch.zeros(1, 20, dtype=torch.long))   # the reference adds all 77 position rows (LD.py:4410)


def test_clip_encode_token_weights_matches_reference_golden(clip_eng):
    """SDClipModel.encode_token_weights with clip skip -2: weighted prompt, two sections, and layer = last."""
    from lightdiffusion_b200.clip import SD1ClipEncoder
    g, sd, eng = clip_eng
    enc = SD1ClipEncoder(eng)
    enc.set_clip_options({"layer": -2})
    cw, pw = enc.encode_token_weights([g["secs"][1]])
    c2, p2 = enc.encode_token_weights([g["secs"][0], g["secs"][2]])
    enc.reset_clip_options()
    cl, _ = enc.encode_token_weights([g["secs"][0]])
    errs = [rel(cw.cpu(), g["cond_w"]), rel(pw.cpu(), g["pooled_w"]), rel(c2.cpu(), g["cond_2"]), rel(p2.cpu(), g["pooled_2"]),
            rel(cl.cpu(), g["cond_last"])]
    print("clip token weights vs reference golden", errs)
    assert cw.shape == (1, 77, 768) and c2.shape == (1, 154, 768) and pw.shape == (1, 768)
    assert max(errs) < COND_TOL


# ------------------------------------------------------------------------------------ ESRGAN upscaler
def test_esrgan_matches_reference_golden(golden):
    """RRDBNet outputs of the real reference (fp32 weights) vs the engine (fp16 weights and activations)."""
    from lightdiffusion_b200.esrgan import ESRGANEngine, esrgan_param_shapes
    g = golden("esrgan.pt")
    for name, c in g["cases"].items():
        sd = sd15.seeded_state_dict(esrgan_param_shapes(c["num_blocks"], c["scale"], arch=c["arch"]), c["seed"], **g["init"])
        eng = ESRGANEngine({k: v.to(DEV) for k, v in sd.items()})
        assert eng.scale == c["scale"] and eng.num_blocks == c["num_blocks"]
        y = eng(c["x"].to(DEV))
        print("esrgan", name, "rel", rel(y.cpu(), c["y"]), "psnr", psnr(y.cpu(), c["y"]))
        assert y.shape == c["y"].shape and y.dtype == torch.float32
        assert rel(y.cpu(), c["y"]) < 5e-3
    with pytest.raises(Exception):
        eng(c["x"])                                   # CPU tensor: no fallback
    with pytest.raises(Exception):
        ESRGANEngine({"foo.weight": torch.zeros(1)})


@pytest.mark.parametrize("b,h,w", [(1, 64, 64), (2, 40, 72), (1, 256, 256)])
def test_esrgan_parity_full_depth(b, h, w):
    """23-block 4x model (the Real-ESRGAN x4 shape) against the oracle on fp16-rounded weights; 256x256 is a tile of
    the reference's tiled_scale at half size."""
    from lightdiffusion_b200.esrgan import ESRGANEngine, esrgan_param_shapes
    sd = rounded(sd15.seeded_state_dict(esrgan_param_shapes(23, 4), 303, bias_std=0.05))
    eng = ESRGANEngine(sd)
    x = torch.rand(b, 3, h, w, generator=torch.Generator().manual_seed(9)).to(DEV)
    y = eng(x)
    with torch.inference_mode():
        ref = O.esrgan_forward(sd, x)
    print("esrgan full depth", (b, h, w), "rel", rel(y, ref))
    assert y.shape == (b, 3, 4 * h, 4 * w)
    assert rel(y, ref) < 5e-3


# ------------------------------------------------------------------------------------ TAESD previews
def test_taesd_matches_reference_golden(golden):
    from lightdiffusion_b200.taesd import TAESDEngine, taesd_param_shapes
    g = golden("taesd.pt")
    dsd = sd15.seeded_state_dict(taesd_param_shapes("decoder"), g["seeds"][0], **g["init"])
    esd = sd15.seeded_state_dict(taesd_param_shapes("encoder"), g["seeds"][1], **g["init"])
    eng = TAESDEngine({k: v.to(DEV) for k, v in dsd.items()}, {k: v.to(DEV) for k, v in esd.items()})
    y = eng.decode(g["latent"].to(DEV))
    z = eng.encode(g["image"].to(DEV))
    print("taesd vs reference golden: decode", rel(y.cpu(), g["decoded"]), "encode", rel(z.cpu(), g["encoded"]))
    assert y.shape == g["decoded"].shape and z.shape == g["encoded"].shape
    assert rel(y.cpu(), g["decoded"]) < 5e-3 and rel(z.cpu(), g["encoded"]) < 5e-3
    # the 64x64 latent of a 512x512 image, against the oracle on fp16-rounded weights
    dr = rounded(dsd)
    lat = (torch.randn(3, 4, 64, 64, generator=torch.Generator().manual_seed(2)) * 2).to(DEV)
    eng2 = TAESDEngine(dr)
    with torch.inference_mode():
        ref = O.taesd_decode(dr, lat)
    assert rel(eng2.decode(lat), ref) < 5e-3
    with pytest.raises(Exception):
        eng2.encode(torch.zeros(1, 3, 64, 64, device=DEV))      # built without encoder weights
    with pytest.raises(Exception):
        eng.decode(g["latent"])                                   # CPU tensor
