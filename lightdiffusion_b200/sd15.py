"""SD1.5 architecture description: block layout, parameter names/shapes, seeded weights.

The reference infers the UNet config from checkpoint keys (LD.py:6065-6182) and builds
``UNetModel1`` from it (LD.py:5294-5686); the VAE config is literal (LD.py:6300-6326).
This module re-derives the same layout as plain data so that weights can be created,
packed and executed without any reference code present (``/root/reference`` does not
exist on the GPU box).  The names equal the reference ``state_dict`` keys so reference
checkpoints load unchanged; ``oracle/make_golden.py`` pins that by a strict
``load_state_dict`` into the real reference modules.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import torch


@dataclass(frozen=True)
class UNetConfig:
    """Subset of the reference ``unet_config`` that shapes the SD1.x UNet (LD.py:5964-5976)."""
    in_channels: int = 4
    out_channels: int = 4
    model_channels: int = 320
    channel_mult: Tuple[int, ...] = (1, 2, 4, 4)
    num_res_blocks: Tuple[int, ...] = (2, 2, 2, 2)
    transformer_depth: Tuple[int, ...] = (1, 1, 1, 1, 1, 1, 0, 0)
    transformer_depth_output: Tuple[int, ...] = (1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0)
    transformer_depth_middle: int = 1
    context_dim: int = 768
    num_heads: int = 8

    @property
    def time_embed_dim(self) -> int:
        return self.model_channels * 4


SD15_UNET = UNetConfig()


def tiny_unet_config(model_channels: int = 64, context_dim: int = 96) -> UNetConfig:
    """Same topology as SD1.5 with fewer channels (fast CPU config)."""
    return UNetConfig(model_channels=model_channels, context_dim=context_dim)


# --------------------------------------------------------------------------- UNet layout
@dataclass
class ResSpec:
    name: str
    cin: int
    cout: int


@dataclass
class AttnSpec:
    name: str
    ch: int
    heads: int
    dim_head: int
    context_dim: int


@dataclass
class ConvSpec:
    name: str          # parameter prefix of the conv itself
    cin: int
    cout: int
    kind: str          # "conv_in" | "down" | "up"


@dataclass
class UNetLayout:
    """Ordered description of the 25 UNet blocks (LD.py:5393-5686)."""
    input_blocks: List[list] = field(default_factory=list)
    middle: list = field(default_factory=list)
    output_blocks: List[list] = field(default_factory=list)
    input_chans: List[int] = field(default_factory=list)


def unet_layout(cfg: UNetConfig) -> UNetLayout:
    mc = cfg.model_channels
    lay = UNetLayout()
    lay.input_blocks.append([ConvSpec("input_blocks.0.0", cfg.in_channels, mc, "conv_in")])
    chans = [mc]
    ch = mc
    td = list(cfg.transformer_depth)
    idx = 1
    nlev = len(cfg.channel_mult)
    for level, mult in enumerate(cfg.channel_mult):
        for _ in range(cfg.num_res_blocks[level]):
            layers = [ResSpec(f"input_blocks.{idx}.0", ch, mult * mc)]
            ch = mult * mc
            if td.pop(0) > 0:
                layers.append(AttnSpec(f"input_blocks.{idx}.1", ch, cfg.num_heads,
                                       ch // cfg.num_heads, cfg.context_dim))
            lay.input_blocks.append(layers)
            chans.append(ch)
            idx += 1
        if level != nlev - 1:
            lay.input_blocks.append([ConvSpec(f"input_blocks.{idx}.0.op", ch, ch, "down")])
            chans.append(ch)
            idx += 1
    lay.input_chans = list(chans)
    lay.middle = [ResSpec("middle_block.0", ch, ch)]
    if cfg.transformer_depth_middle >= 0:
        lay.middle += [AttnSpec("middle_block.1", ch, cfg.num_heads, ch // cfg.num_heads,
                                cfg.context_dim),
                       ResSpec("middle_block.2", ch, ch)]
    tdo = list(cfg.transformer_depth_output)
    oidx = 0
    for level, mult in list(enumerate(cfg.channel_mult))[::-1]:
        for i in range(cfg.num_res_blocks[level] + 1):
            ich = chans.pop()
            layers = [ResSpec(f"output_blocks.{oidx}.0", ch + ich, mc * mult)]
            ch = mc * mult
            if tdo.pop() > 0:
                layers.append(AttnSpec(f"output_blocks.{oidx}.1", ch, cfg.num_heads,
                                       ch // cfg.num_heads, cfg.context_dim))
            if level and i == cfg.num_res_blocks[level]:
                layers.append(ConvSpec(f"output_blocks.{oidx}.{len(layers)}.conv", ch, ch, "up"))
            lay.output_blocks.append(layers)
            oidx += 1
    return lay


def _res_params(out: Dict[str, tuple], s: ResSpec, temb: int) -> None:
    p = s.name
    out[f"{p}.in_layers.0.weight"] = (s.cin,)
    out[f"{p}.in_layers.0.bias"] = (s.cin,)
    out[f"{p}.in_layers.2.weight"] = (s.cout, s.cin, 3, 3)
    out[f"{p}.in_layers.2.bias"] = (s.cout,)
    out[f"{p}.emb_layers.1.weight"] = (s.cout, temb)
    out[f"{p}.emb_layers.1.bias"] = (s.cout,)
    out[f"{p}.out_layers.0.weight"] = (s.cout,)
    out[f"{p}.out_layers.0.bias"] = (s.cout,)
    out[f"{p}.out_layers.3.weight"] = (s.cout, s.cout, 3, 3)
    out[f"{p}.out_layers.3.bias"] = (s.cout,)
    if s.cin != s.cout:
        out[f"{p}.skip_connection.weight"] = (s.cout, s.cin, 1, 1)
        out[f"{p}.skip_connection.bias"] = (s.cout,)


def _attn_params(out: Dict[str, tuple], s: AttnSpec) -> None:
    p, c = s.name, s.ch
    inner = s.heads * s.dim_head
    out[f"{p}.norm.weight"] = (c,)
    out[f"{p}.norm.bias"] = (c,)
    out[f"{p}.proj_in.weight"] = (inner, c, 1, 1)
    out[f"{p}.proj_in.bias"] = (inner,)
    t = f"{p}.transformer_blocks.0"
    out[f"{t}.attn1.to_q.weight"] = (inner, inner)
    out[f"{t}.attn1.to_k.weight"] = (inner, inner)
    out[f"{t}.attn1.to_v.weight"] = (inner, inner)
    out[f"{t}.attn1.to_out.0.weight"] = (inner, inner)
    out[f"{t}.attn1.to_out.0.bias"] = (inner,)
    out[f"{t}.ff.net.0.proj.weight"] = (inner * 8, inner)
    out[f"{t}.ff.net.0.proj.bias"] = (inner * 8,)
    out[f"{t}.ff.net.2.weight"] = (inner, inner * 4)
    out[f"{t}.ff.net.2.bias"] = (inner,)
    out[f"{t}.attn2.to_q.weight"] = (inner, inner)
    out[f"{t}.attn2.to_k.weight"] = (inner, s.context_dim)
    out[f"{t}.attn2.to_v.weight"] = (inner, s.context_dim)
    out[f"{t}.attn2.to_out.0.weight"] = (inner, inner)
    out[f"{t}.attn2.to_out.0.bias"] = (inner,)
    for n in ("norm1", "norm2", "norm3"):
        out[f"{t}.{n}.weight"] = (inner,)
        out[f"{t}.{n}.bias"] = (inner,)
    out[f"{p}.proj_out.weight"] = (c, inner, 1, 1)
    out[f"{p}.proj_out.bias"] = (c,)


def unet_param_shapes(cfg: UNetConfig = SD15_UNET) -> "OrderedDict[str, tuple]":
    """name -> shape for every UNet parameter, keys equal to the reference state_dict."""
    lay = unet_layout(cfg)
    out: "OrderedDict[str, tuple]" = OrderedDict()
    mc, te = cfg.model_channels, cfg.time_embed_dim
    out["time_embed.0.weight"] = (te, mc)
    out["time_embed.0.bias"] = (te,)
    out["time_embed.2.weight"] = (te, te)
    out["time_embed.2.bias"] = (te,)
    for block in lay.input_blocks + [lay.middle] + lay.output_blocks:
        for s in block:
            if isinstance(s, ResSpec):
                _res_params(out, s, te)
            elif isinstance(s, AttnSpec):
                _attn_params(out, s)
            else:
                out[f"{s.name}.weight"] = (s.cout, s.cin, 3, 3)
                out[f"{s.name}.bias"] = (s.cout,)
    out["out.0.weight"] = (mc,)
    out["out.0.bias"] = (mc,)
    out["out.2.weight"] = (cfg.out_channels, mc, 3, 3)
    out["out.2.bias"] = (cfg.out_channels,)
    return out


# --------------------------------------------------------------------------- VAE decoder
@dataclass(frozen=True)
class VAEConfig:
    """The literal AutoencoderKL config of LD.py:6300-6326 (decoder half)."""
    z_channels: int = 4
    out_ch: int = 3
    in_channels: int = 3
    ch: int = 128
    ch_mult: Tuple[int, ...] = (1, 2, 4, 4)
    num_res_blocks: int = 2


SD15_VAE = VAEConfig()


def _vae_res(out, p, cin, cout):
    out[f"{p}.norm1.weight"] = (cin,)
    out[f"{p}.norm1.bias"] = (cin,)
    out[f"{p}.conv1.weight"] = (cout, cin, 3, 3)
    out[f"{p}.conv1.bias"] = (cout,)
    out[f"{p}.norm2.weight"] = (cout,)
    out[f"{p}.norm2.bias"] = (cout,)
    out[f"{p}.conv2.weight"] = (cout, cout, 3, 3)
    out[f"{p}.conv2.bias"] = (cout,)
    if cin != cout:
        out[f"{p}.nin_shortcut.weight"] = (cout, cin, 1, 1)
        out[f"{p}.nin_shortcut.bias"] = (cout,)


def vae_decoder_levels(cfg: VAEConfig = SD15_VAE):
    """[(i_level, [(cin,cout)...], has_upsample)] in execution order (LD.py:3822-3847)."""
    n = len(cfg.ch_mult)
    block_in = cfg.ch * cfg.ch_mult[n - 1]
    levels = []
    for i_level in reversed(range(n)):
        block_out = cfg.ch * cfg.ch_mult[i_level]
        blocks = []
        for _ in range(cfg.num_res_blocks + 1):
            blocks.append((block_in, block_out))
            block_in = block_out
        levels.append((i_level, blocks, i_level != 0))
    return levels


def vae_decoder_param_shapes(cfg: VAEConfig = SD15_VAE) -> "OrderedDict[str, tuple]":
    """name -> shape for ``post_quant_conv`` + ``decoder.*`` (keys of AutoencodingEngine)."""
    out: "OrderedDict[str, tuple]" = OrderedDict()
    out["post_quant_conv.weight"] = (cfg.z_channels, cfg.z_channels, 1, 1)
    out["post_quant_conv.bias"] = (cfg.z_channels,)
    top = cfg.ch * cfg.ch_mult[-1]
    out["decoder.conv_in.weight"] = (top, cfg.z_channels, 3, 3)
    out["decoder.conv_in.bias"] = (top,)
    _vae_res(out, "decoder.mid.block_1", top, top)
    out["decoder.mid.attn_1.norm.weight"] = (top,)
    out["decoder.mid.attn_1.norm.bias"] = (top,)
    for n in ("q", "k", "v", "proj_out"):
        out[f"decoder.mid.attn_1.{n}.weight"] = (top, top, 1, 1)
        out[f"decoder.mid.attn_1.{n}.bias"] = (top,)
    _vae_res(out, "decoder.mid.block_2", top, top)
    for i_level, blocks, has_up in vae_decoder_levels(cfg):
        for j, (cin, cout) in enumerate(blocks):
            _vae_res(out, f"decoder.up.{i_level}.block.{j}", cin, cout)
        if has_up:
            c = blocks[-1][1]
            out[f"decoder.up.{i_level}.upsample.conv.weight"] = (c, c, 3, 3)
            out[f"decoder.up.{i_level}.upsample.conv.bias"] = (c,)
    last = cfg.ch * cfg.ch_mult[0]
    out["decoder.norm_out.weight"] = (last,)
    out["decoder.norm_out.bias"] = (last,)
    out["decoder.conv_out.weight"] = (cfg.out_ch, last, 3, 3)
    out["decoder.conv_out.bias"] = (cfg.out_ch,)
    return out


# --------------------------------------------------------------------------- VAE encoder
def vae_encoder_levels(cfg: VAEConfig = SD15_VAE):
    """[(i_level, [(cin,cout)...], has_downsample)] in execution order (Encoder.__init__, LD.py:3684-3708)."""
    n = len(cfg.ch_mult)
    in_mult = (1,) + tuple(cfg.ch_mult)
    levels = []
    for i_level in range(n):
        block_in, block_out = cfg.ch * in_mult[i_level], cfg.ch * cfg.ch_mult[i_level]
        blocks = []
        for _ in range(cfg.num_res_blocks):
            blocks.append((block_in, block_out))
            block_in = block_out
        levels.append((i_level, blocks, i_level != n - 1))
    return levels


def vae_encoder_param_shapes(cfg: VAEConfig = SD15_VAE) -> "OrderedDict[str, tuple]":
    """name -> shape for ``encoder.*`` + ``quant_conv`` (keys of AutoencodingEngine, LD.py:3461-3481;
    double_z: the encoder emits mean and logvar, 2 * z_channels)."""
    out: "OrderedDict[str, tuple]" = OrderedDict()
    out["encoder.conv_in.weight"] = (cfg.ch, cfg.in_channels, 3, 3)
    out["encoder.conv_in.bias"] = (cfg.ch,)
    for i_level, blocks, has_down in vae_encoder_levels(cfg):
        for j, (cin, cout) in enumerate(blocks):
            _vae_res(out, f"encoder.down.{i_level}.block.{j}", cin, cout)
        if has_down:
            c = blocks[-1][1]
            out[f"encoder.down.{i_level}.downsample.conv.weight"] = (c, c, 3, 3)
            out[f"encoder.down.{i_level}.downsample.conv.bias"] = (c,)
    top = cfg.ch * cfg.ch_mult[-1]
    _vae_res(out, "encoder.mid.block_1", top, top)
    out["encoder.mid.attn_1.norm.weight"] = (top,)
    out["encoder.mid.attn_1.norm.bias"] = (top,)
    for n in ("q", "k", "v", "proj_out"):
        out[f"encoder.mid.attn_1.{n}.weight"] = (top, top, 1, 1)
        out[f"encoder.mid.attn_1.{n}.bias"] = (top,)
    _vae_res(out, "encoder.mid.block_2", top, top)
    out["encoder.norm_out.weight"] = (top,)
    out["encoder.norm_out.bias"] = (top,)
    out["encoder.conv_out.weight"] = (2 * cfg.z_channels, top, 3, 3)
    out["encoder.conv_out.bias"] = (2 * cfg.z_channels,)
    out["quant_conv.weight"] = (2 * cfg.z_channels, 2 * cfg.z_channels, 1, 1)
    out["quant_conv.bias"] = (2 * cfg.z_channels,)
    return out


# --------------------------------------------------------------------------- CLIP text tower
@dataclass(frozen=True)
class CLIPTextConfig:
    """_internal/clip/sd1_clip_config.json (CLIP ViT-L/14 text model), read by SDClipModel (LD.py:4596-4602)."""
    vocab_size: int = 49408
    max_positions: int = 77
    hidden_size: int = 768
    num_layers: int = 12
    num_heads: int = 12
    intermediate_size: int = 3072


SD15_CLIP = CLIPTextConfig()


def clip_text_param_shapes(cfg: CLIPTextConfig = SD15_CLIP) -> "OrderedDict[str, tuple]":
    """name -> shape of CLIPTextModel (LD.py:4456-4467: ``text_model.*`` + ``text_projection``)."""
    out: "OrderedDict[str, tuple]" = OrderedDict()
    c, f = cfg.hidden_size, cfg.intermediate_size
    out["text_model.embeddings.token_embedding.weight"] = (cfg.vocab_size, c)
    out["text_model.embeddings.position_embedding.weight"] = (cfg.max_positions, c)
    for i in range(cfg.num_layers):
        p = f"text_model.encoder.layers.{i}"
        out[f"{p}.layer_norm1.weight"] = (c,)
        out[f"{p}.layer_norm1.bias"] = (c,)
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            out[f"{p}.self_attn.{n}.weight"] = (c, c)
            out[f"{p}.self_attn.{n}.bias"] = (c,)
        out[f"{p}.layer_norm2.weight"] = (c,)
        out[f"{p}.layer_norm2.bias"] = (c,)
        out[f"{p}.mlp.fc1.weight"] = (f, c)
        out[f"{p}.mlp.fc1.bias"] = (f,)
        out[f"{p}.mlp.fc2.weight"] = (c, f)
        out[f"{p}.mlp.fc2.bias"] = (c,)
    out["text_model.final_layer_norm.weight"] = (c,)
    out["text_model.final_layer_norm.bias"] = (c,)
    out["text_projection.weight"] = (c, c)
    return out


# --------------------------------------------------------------------------- weights
def seeded_state_dict(shapes: "OrderedDict[str, tuple]", seed: int, dtype=torch.float32,
                      device="cpu", bias_std: float = 0.0, norm_jitter: float = 0.0):
    """Deterministic random-init weights.

    The reference leaves parameters uninitialised (``reset_parameters`` is a no-op,
    LD.py:2363-2399), so the harness defines them: N(0, 1/fan_in) for tensors with
    ndim >= 2 (including the ``zero_module``'d ``out.2``), 1.0 for norm weights, 0 for
    biases.  ``bias_std`` / ``norm_jitter`` > 0 make biases and norm affine parameters
    non-trivial so that parity tests exercise them.
    Values are drawn with a CPU generator in key order, so they are identical on every
    host; ``device='cuda'`` draws on the device instead (bench only, values differ).
    """
    dev = torch.device(device)
    g = torch.Generator(device=dev if dev.type == "cuda" else "cpu").manual_seed(seed)
    gdev = dev if dev.type == "cuda" else "cpu"
    sd = OrderedDict()
    for name, shape in shapes.items():
        if len(shape) >= 2:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            v = torch.randn(shape, generator=g, device=gdev) / fan_in ** 0.5
        elif name.endswith("weight"):
            v = torch.ones(shape, device=gdev)
            if norm_jitter:
                v = v + norm_jitter * torch.randn(shape, generator=g, device=gdev)
        else:
            v = torch.zeros(shape, device=gdev)
            if bias_std:
                v = bias_std * torch.randn(shape, generator=g, device=gdev)
        sd[name] = v.to(device=dev, dtype=dtype)
    return sd


def count_params(shapes) -> int:
    n = 0
    for s in shapes.values():
        k = 1
        for d in s:
            k *= d
        n += k
    return n
