"""Pins the oracle restatement (oracle/ldm_oracle.py) to golden vectors produced by the REAL
reference code (oracle/make_golden.py, run in the build container).  CPU only."""
import torch

from lightdiffusion_b200 import sd15
from oracle import ldm_oracle as O


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def test_schedules(golden):
    g = golden("schedules.pt")
    table = O.make_sigmas_table()
    assert torch.equal(table.sigmas, g["sigmas_table"])
    assert abs(float(table.sigma_min) - 0.029167) < 1e-6 and abs(float(table.sigma_max) - 14.614641) < 1e-5
    assert torch.equal(O.calculate_sigmas("karras", 20, table), g["karras20"])
    assert torch.equal(O.calculate_sigmas("karras", 8, table), g["karras8"])
    assert torch.equal(O.calculate_sigmas("normal", 10, table), g["normal10"])
    assert torch.equal(O.calculate_sigmas("normal", 22, table), g["normal22"])
    assert torch.equal(O.timestep_of_sigma(g["probe_sigmas"], table.log_sigmas), g["probe_timesteps"])
    assert torch.equal(O.sigma_of_timestep(g["probe_t"], table.log_sigmas), g["probe_sigma_of_t"])
    # SURVEY.md §4 known answers
    k = O.calculate_sigmas("karras", 20, table)
    assert abs(float(k[1]) - 11.7254) < 1e-3 and abs(float(k[10]) - 1.0908) < 1e-3 and float(k[-1]) == 0


def test_denoise_schedule_slice(golden):
    g = golden("samplers_tiny.pt")
    s = O.sigmas_for_denoise("normal", 4, 0.45)
    assert torch.equal(s, g["sigmas2"])


def test_samplers_toy(golden):
    g = golden("samplers_toy.pt")

    def model(x, sigma, **kw):
        return x / (1.0 + sigma.view(-1, 1, 1, 1) ** 2)
    a = O.sample_dpmpp_2m(model, g["x0"].clone(), g["sigmas"])
    assert torch.equal(a, g["dpmpp2m"])
    torch.manual_seed(g["euler_global_seed"])
    b = O.sample_euler_ancestral(model, g["x0"].clone(), g["sigmas"])
    assert torch.equal(b, g["euler_a"])


def _check_unet(g, cfg):
    sd = sd15.seeded_state_dict(sd15.unet_param_shapes(cfg), g["unet_seed"], **g["init"])
    assert sd15.count_params(sd15.unet_param_shapes(cfg)) == g["n_params"]
    with torch.inference_mode():
        eps = O.unet_forward(sd, cfg, g["x"] / (g["sigma"].view(-1, 1, 1, 1) ** 2 + 1) ** 0.5,
                             g["timestep"], g["ctx"])
        assert rel(eps, g["eps"]) < 2e-5
        den = O.apply_model(sd, cfg, g["x"], g["sigma"], g["ctx"])
        assert rel(den, g["denoised"]) < 2e-5
        model = O.cfg_denoiser(sd, cfg, g["pos"], g["neg"], g["cfg_scale"])
        assert rel(model(g["x"], g["sigma"]), g["cfg_denoised"]) < 2e-5
    return sd


def test_unet_tiny(golden):
    g = golden("unet_tiny.pt")
    _check_unet(g, g["cfg"])


def test_unet_sd15_dims(golden):
    g = golden("unet_sd15_lat8.pt")
    assert g["n_params"] == 859520964                  # 859.521 M
    _check_unet(g, sd15.SD15_UNET)


def test_samplers_tiny_model(golden):
    gu, g = golden("unet_tiny.pt"), golden("samplers_tiny.pt")
    cfg = gu["cfg"]
    sd = sd15.seeded_state_dict(sd15.unet_param_shapes(cfg), gu["unet_seed"], **gu["init"])
    with torch.inference_mode():
        z = O.sample_latent(sd, cfg, g["noise"], g["pos"], g["neg"], g["cfg_scale"], g["sigmas"], "dpmpp_2m")
        assert rel(z, g["z_dpmpp2m"]) < 1e-4
        torch.manual_seed(g["euler_global_seed"])
        z2 = O.sample_latent(sd, cfg, g["noise2"], g["pos"], g["neg"], g["cfg_scale2"], g["sigmas2"],
                             "euler_ancestral", latent=g["lat2"])
        assert rel(z2, g["z_euler_a"]) < 1e-4


def test_vae_decode(golden):
    g = golden("vae_decode.pt")
    shapes = sd15.vae_decoder_param_shapes()
    assert sum(v for k, v in ((k, torch.Size(s).numel()) for k, s in shapes.items())
               if k.startswith("decoder.")) == g["n_decoder_params"]
    sd = sd15.seeded_state_dict(shapes, g["vae_seed"], **g["init"])
    with torch.inference_mode():
        img = O.vae_decode(sd, sd15.SD15_VAE, g["z"])
    assert img.shape == g["image"].shape == (2, 64, 64, 3)
    assert rel(img, g["image"]) < 2e-5


def test_vae_encode(golden):
    """The encoder restatement against the reference's VAE.encode output (fixture from oracle/make_golden.py)."""
    g = golden("vae_encode.pt")
    shapes = sd15.vae_encoder_param_shapes()
    assert sum(torch.Size(s).numel() for k, s in shapes.items() if k.startswith("encoder.")) == g["n_encoder_params"]
    sd = sd15.seeded_state_dict(shapes, g["vae_seed"], **g["init"])
    with torch.inference_mode():
        mom = O.vae_encode_moments(sd, sd15.SD15_VAE, g["pixels"].movedim(-1, 1) * 2.0 - 1.0)
        z = O.vae_encode(sd, sd15.SD15_VAE, g["pixels"], g["noise"])
    assert mom.shape == g["moments"].shape == (2, 8, 4, 6)
    assert rel(mom, g["moments"]) < 2e-5
    assert z.shape == g["z"].shape == (2, 4, 4, 6)
    assert rel(z, g["z"]) < 2e-5
    # the reference draws the posterior noise from the global CPU generator (LD.py:175-179)
    torch.manual_seed(g["global_seed"])
    assert torch.equal(torch.randn(2, 4, 4, 6), g["noise"])


def test_bislerp(golden):
    """The bislerp restatement against the reference's own outputs, edge cases included."""
    g = golden("bislerp.pt")
    for name, c in g["cases"].items():
        x = c.get("x", g["x"])
        out = O.bislerp(x, c["width"], c["height"])
        assert out.shape == c["out"].shape, name
        assert torch.allclose(out, c["out"], rtol=1e-5, atol=1e-6), name


def test_clip_text(golden):
    """The CLIP text tower restatement against the reference's SDClipModel (raw forward, clip-skip -2, token weights)."""
    g = golden("clip_text.pt")
    cfg = sd15.SD15_CLIP
    sd = sd15.seeded_state_dict(sd15.clip_text_param_shapes(cfg), g["clip_seed"], **g["init"])
    with torch.inference_mode():
        last, inter, pooled_proj, pooled = O.clip_text_forward(sd, cfg, g["tokens"], -2)
        assert rel(last, g["last"]) < 2e-5 and rel(inter, g["inter"]) < 2e-5
        assert rel(pooled, g["pooled"]) < 2e-5 and rel(pooled_proj, g["pooled_proj"]) < 2e-5
        assert O.clip_text_forward(sd, cfg, g["tokens"][:1])[1] is None
        assert rel(O.clip_text_forward(sd, cfg, g["tokens"][:1])[0], g["last_only"]) < 2e-5
        cw, pw = O.clip_encode_token_weights(sd, cfg, [g["secs"][1]], -2, g["empty"])
        assert cw.shape == g["cond_w"].shape == (1, 77, 768)
        assert rel(cw, g["cond_w"]) < 2e-5 and rel(pw, g["pooled_w"]) < 2e-5
        c2, p2 = O.clip_encode_token_weights(sd, cfg, [g["secs"][0], g["secs"][2]], -2, g["empty"])
        assert c2.shape == g["cond_2"].shape == (1, 154, 768)
        assert rel(c2, g["cond_2"]) < 2e-5 and rel(p2, g["pooled_2"]) < 2e-5
        cl, _ = O.clip_encode_token_weights(sd, cfg, [g["secs"][0]], None, g["empty"])
        assert rel(cl, g["cond_last"]) < 2e-5


DPM_MODELS = {"linear": lambda x, sigma, **kw: x / (1.0 + sigma.view(-1, 1, 1, 1) ** 2),
              "tanh": lambda x, sigma, **kw: torch.tanh(x / (1.0 + sigma.view(-1, 1, 1, 1))) * 1.5}


def test_dpm_adaptive(golden):
    """The DPM-Solver-12/23 restatement against the reference's sample_dpm_adaptive: same accepted / rejected step
    counts (default, tight-tolerance and PID variants) and the same final state."""
    g = golden("dpm_adaptive.pt")
    for (name, opts), c in g["cases"].items():
        x, info = O.sample_dpm_adaptive(DPM_MODELS[name], g["x0"].clone(), g["sigma_min"], g["sigma_max"],
                                        return_info=True, **dict(opts))
        assert info == c["info"], (name, opts, info, c["info"])
        assert rel(x, c["x"]) < 2e-6, (name, opts)


def test_esrgan(golden):
    """The RRDBNet restatement against the reference's model built by its own loader from both checkpoint layouts."""
    from lightdiffusion_b200.esrgan import esrgan_param_shapes, to_old_arch
    g = golden("esrgan.pt")
    for name, c in g["cases"].items():
        sd = sd15.seeded_state_dict(esrgan_param_shapes(c["num_blocks"], c["scale"], arch=c["arch"]), c["seed"], **g["init"])
        with torch.inference_mode():
            y = O.esrgan_forward(sd, c["x"])
        assert y.shape == c["y"].shape == (2, 3, c["x"].shape[2] * c["scale"], c["x"].shape[3] * c["scale"])
        assert rel(y, c["y"]) < 2e-5, name
        # the three layouts name the same tensors
        old = to_old_arch(sd)
        assert list(old) == list(esrgan_param_shapes(c["num_blocks"], c["scale"], arch="old")) or \
            set(old) == set(esrgan_param_shapes(c["num_blocks"], c["scale"], arch="old"))


def test_taesd(golden):
    """The TAESD restatement against the reference's Decoder2 / Encoder2 stacks."""
    from lightdiffusion_b200.taesd import taesd_param_shapes
    g = golden("taesd.pt")
    dsd = sd15.seeded_state_dict(taesd_param_shapes("decoder"), g["seeds"][0], **g["init"])
    esd = sd15.seeded_state_dict(taesd_param_shapes("encoder"), g["seeds"][1], **g["init"])
    with torch.inference_mode():
        assert rel(O.taesd_decode(dsd, g["latent"]), g["decoded"]) < 2e-5
        assert rel(O.taesd_encode(esd, g["image"]), g["encoded"]) < 2e-5
    assert g["decoded"].shape == (2, 3, 64, 96) and g["encoded"].shape == (2, 4, 8, 6)
