"""ORACLE — test infrastructure, not product code.

Loads the *real* reference (/root/reference/LightDiffusion.py) into a module object so
that ``make_golden.py`` can run the reference's own classes and dump golden vectors.
The recipe is SURVEY.md Appendix B: stub the GUI / detector imports, skip the six
hf_hub_download blocks (LD.py:53-120), force the CPU state on a GPU-less host
(LD.py:1380) and inject a stub ``app`` (LD.py:922-927).

Where the reference comes from: $LDX_REFERENCE_DIR, default /root/reference (the build container).  The GPU box
has no /root/reference; there the unmodified copy that travels with the repo under baseline/_ref/reference
(git-ignored, not gpurun-ignored) is used -- by tests/test_dropin_gpu.py (the reference drives the ldx seams) and by
bench.py's reference legs (the reference's own CPU path, and its PyTorch fp16 path on the same GPU).  Both set
LDX_REFERENCE_DIR before importing this module and skip / fall back to the oracle port when the copy is absent.
Product code never imports this file.
"""
from __future__ import annotations

import glob as _glob
import os
import sys
import types
from unittest import mock

import torch

REF_DIR = os.environ.get("LDX_REFERENCE_DIR", "/root/reference")


def reference_available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "LightDiffusion.py"))


_CACHED = {}


def load_reference(force_cpu: bool = False):
    """exec the reference source into a module (≈16 s) and return it.  ``force_cpu`` loads a second, independent
    copy whose device manager is pinned to the CPU state (LD.py:1380) even on a GPU box: the reference's own CPU
    path, timed by bench.py beside the GPU numbers."""
    cpu = force_cpu or not torch.cuda.is_available()
    if cpu in _CACHED:
        return _CACHED[cpu]
    for name in ["tkinter", "tkinter.filedialog", "customtkinter", "ollama", "torchsde",
                 "segment_anything", "ultralytics", "PIL.ImageTk"]:
        m = mock.MagicMock(name=name)
        m.__all__ = []
        sys.modules[name] = m
    import tkinter
    tkinter.Tk = type("Tk", (object,), {})
    import PIL
    PIL.ImageTk = sys.modules["PIL.ImageTk"]
    path = os.path.join(REF_DIR, "LightDiffusion.py")
    src = open(path).read()
    if cpu:
        assert src.count("cpu_state = CPUState.GPU\n") == 1
        src = src.replace("cpu_state = CPUState.GPU\n", "cpu_state = CPUState.CPU\n")
    real_glob = _glob.glob
    _glob.glob = lambda p, *a, **k: ["stub"] if str(p).startswith("./_internal/") else real_glob(p, *a, **k)
    cwd = os.getcwd()
    os.chdir(REF_DIR)
    try:
        name = "LightDiffusion_ref_cpu" if (cpu and torch.cuda.is_available()) else "LightDiffusion_ref"
        mod = types.ModuleType(name)
        mod.__file__ = path
        sys.modules[name] = mod
        exec(compile(src, path, "exec"), mod.__dict__)
    finally:
        os.chdir(cwd)
        _glob.glob = real_glob

    class _CB:
        get = staticmethod(lambda: False)

    class _App:
        interrupt_flag = False
        previewer_checkbox = _CB()
        title = staticmethod(lambda *a: None)

    mod.app = _App()
    mod.PROGRESS_BAR_ENABLED = False
    _CACHED[cpu] = mod
    return mod


def ref_unet_config(cfg) -> dict:
    """The dict detect_unet_config (LD.py:6065-6182) yields for SD1.x, from our UNetConfig."""
    return dict(use_checkpoint=False, image_size=32, use_spatial_transformer=True, legacy=False,
                adm_in_channels=None, in_channels=cfg.in_channels, out_channels=cfg.out_channels,
                model_channels=cfg.model_channels, num_res_blocks=list(cfg.num_res_blocks),
                channel_mult=list(cfg.channel_mult), transformer_depth=list(cfg.transformer_depth),
                transformer_depth_output=list(cfg.transformer_depth_output),
                transformer_depth_middle=cfg.transformer_depth_middle, use_linear_in_transformer=False,
                context_dim=cfg.context_dim, use_temporal_resblock=False, use_temporal_attention=False)


def build_ref_model(R, cfg, sd, dtype=torch.float32):
    """Reference BaseModel (LD.py:5798) with our seeded weights loaded strictly."""
    ucfg = ref_unet_config(cfg)
    mc = R.model_config_from_unet_config(ucfg)       # sm_SD15 for the real SD1.5 dims
    if mc is None:                                    # reduced-width test config: same class, built directly
        mc = R.sm_SD15(ucfg)
    assert isinstance(mc, R.sm_SD15) and mc.unet_config["num_heads"] == cfg.num_heads
    mc.set_inference_dtype(dtype, None)
    model = R.BaseModel(mc, device="cpu")
    model.diffusion_model.load_state_dict({k: v.to(dtype) for k, v in sd.items()}, strict=True)
    return model.eval()


def build_ref_vae(R, sd, dtype=torch.float32, device=None):
    """Reference VAE (LD.py:6282) with our seeded decoder and/or encoder weights (the other half left as built)."""
    vae = R.VAE(sd={}, device=torch.device("cpu") if device is None else device, dtype=dtype)
    fsm = vae.first_stage_model
    own = fsm.state_dict()
    halves = (("encoder.", "quant_conv.") if any(k.startswith("encoder.") for k in sd) else ()) + \
             (("decoder.", "post_quant_conv.") if any(k.startswith("decoder.") for k in sd) else ())
    missing = [k for k in own if k.startswith(halves) and k not in sd]
    extra = [k for k in sd if k not in own]
    assert not missing and not extra, (missing, extra)
    with torch.no_grad():
        for k, v in sd.items():
            assert tuple(own[k].shape) == tuple(v.shape), k
            own[k].copy_(v.to(dtype))
    return vae
