"""Whole-network execution plans for the SD1.5 UNet (UNetModel1.forward, LD.py:5688-5767, through
BaseModel.apply_model, LD.py:5828-5860) and the VAE decoder (Decoder.forward LD.py:3857-3882 through
VAE.decode LD.py:6357-6381), built from a reference-format state_dict.

A plan is a static list of libldx.so kernel launches over pre-allocated NHWC fp16 buffers for one
(batch, height, width); it is replayed through ldx_program_run and captured in a CUDA graph.
Host code here only wires pointers: every FLOP runs in the hand-written kernels, and construction
fails loudly when libldx.so or a sm_100 device is missing.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import os

import torch

from . import _lib as L
from . import kernels as K
from .sd15 import (SD15_UNET, SD15_VAE, AttnSpec, ConvSpec, ResSpec, UNetConfig, VAEConfig, unet_layout,
                   vae_decoder_levels, vae_encoder_levels)

F16, F32 = torch.float16, torch.float32


def sigma_tables():
    """The 1000 discrete sigmas / log-sigmas of ModelSamplingDiscrete (LD.py:1280-1326): scaled-linear
    betas in float64; log taken in float64 then cast (bit-exact with the reference buffers)."""
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float64) ** 2
    acp = torch.cumprod(1.0 - betas, dim=0)
    sig = ((1 - acp) / acp) ** 0.5
    return sig.float(), sig.log().float()


# norm3 feeds the GEGLU projection, whose epilogue (erf-GELU on every gate) is already the longest in the
# UNet; folding the LayerNorm in as well is a small net win (17.27 -> 17.19 ms per UNet step on B200).
# LDX_FUSE_LN3=0 keeps the separate LayerNorm kernel (A/B measurements).
FUSE_LN3 = os.environ.get("LDX_FUSE_LN3", "1") == "1"
# Under classifier-free guidance both halves of the [uncond, cond] batch carry the same latent and timestep
# (LD.py:2545-2551); the layers before the first cross-attention therefore run once per latent (bit-identical
# results).  LDX_SHARE_CFG_PREFIX=0 evaluates the full 2b batch everywhere.
SHARE_CFG_PREFIX = os.environ.get("LDX_SHARE_CFG_PREFIX", "1") == "1"
# A plan owns its activation pool (~0.5 GB for the SD1.5 UNet at B = 16) and a CUDA graph; an engine keeps the most
# recently used MAX_PLANS of them (an img2img flow over varied tile sizes would otherwise exhaust HBM).
MAX_PLANS = int(os.environ.get("LDX_MAX_PLANS", "4"))
# Convolutions with <= 5 output channels (UNet out 320 -> 4, VAE conv_out 128 -> 3) run on the dedicated
# one-read kernel (ldx_conv3x3_small); LDX_CONV_SMALL=0 sends them through the implicit-GEMM kernel (A/B).
CONV_SMALL = os.environ.get("LDX_CONV_SMALL", "1") == "1"
# GroupNorm statistics come out of the producing GEMM's epilogue (column partials per 32-row slot) instead of a
# statistics pass over the tensor; LDX_FUSE_GN_STATS=0 keeps the two-pass GroupNorm everywhere (A/B).
# Measured on B200 (graph replay): the VAE decoder gains (512^2 b=8: 21.3 -> 20.2 ms, 1024^2 b=4: 45.7 -> 44.0 ms: its
# statistics passes read 0.27-0.54 GB maps from HBM), the SD1.5 UNet at 64x64 does not (15.54 -> 15.65 ms: its 21-42 MB
# maps are still in L2 when the 10 us statistics pass reads them, and the epilogue work costs the GEMMs as much as it
# saves).  Hence the threshold: maps of at least GN_STATS_MIN_HW pixels per sample (LDX_GN_STATS_MIN_HW overrides).
# ResBlock convs (GroupNorm -> SiLU -> 3x3 conv, LD.py:5220-5261 / 3562-3576) on maps of at least HALO_MIN_HW pixels
# run as the halo-tile conv with the normalisation applied in its operand path (csrc/conv_halo.cu): the GroupNorm
# becomes a statistics-only launch (ldx_group_norm_coef) and its apply pass (one read + one write of the map)
# disappears.  LDX_HALO=0 keeps the two-pass GroupNorm + implicit-GEMM conv everywhere (A/B).
HALO = os.environ.get("LDX_HALO", "1") == "1"
HALO_MIN_HW = int(os.environ.get("LDX_HALO_MIN_HW", "256"))
HALO_MIN_TILES = int(os.environ.get("LDX_HALO_MIN_TILES", "32"))
# SpatialTransformer.norm folded into proj_in's operand path the same way (a 1x1 conv through the halo kernel).  Measured
# slower on B200 (UNet step +0.17 ms: with one MMA chain of 5-20 K steps per item instead of nine, the three-deep
# operand ring cannot cover TMA latency + transform), so it is off by default; LDX_HALO_1X1=1 enables it.
HALO_1X1 = os.environ.get("LDX_HALO_1X1", "0") == "1"
# The feed-forward of the 320-channel transformer blocks (GEGLU projection + second Linear + residual, LD.py:4155-4162) as
# ONE kernel whose 4C-wide intermediate stays on chip (csrc/ff_fused.cu).  LDX_FF_FUSED=0 keeps the two GEMMs (A/B).
FF_FUSED = os.environ.get("LDX_FF_FUSED", "1") == "1"
# SpatialTransformer.norm folded into proj_in's weights per sample (ldx_gn_fold_weights + ldx_gemm_desc_t.w_batch_rows) where the
# map is large enough for the two-pass GroupNorm (the 64x64 and 32x32 levels): the normalised map is never written or read.
# Correct (tests/test_kernels_gpu.py::test_group_norm_folded_into_proj_in_weights, UNet eps 1.50e-3 vs 1.48e-3) but measured
# SLOWER and therefore off: the ten weight-scaling launches cost ~19 us each against the 9 us of the gn_apply pass they
# replace (graph replay 14.97 / 14.98 vs 14.80 / 14.96 ms; bench conditions 24.39 / 24.38 vs 24.58 / 24.58 images/s,
# scripts/gpu_ab10.sh).
GN_FOLD = os.environ.get("LDX_GN_FOLD", "0") == "1"
FUSE_GN_STATS = os.environ.get("LDX_FUSE_GN_STATS", "1") == "1"
GN_STATS_MIN_HW = int(os.environ.get("LDX_GN_STATS_MIN_HW", "16384"))


class _PlanCache:
    """LRU map key -> plan.  Evicted plans are dropped (their pool, graph and recorded program are freed)."""

    def __init__(self, capacity: int = MAX_PLANS):
        self.capacity, self._d = max(1, capacity), OrderedDict()

    def get(self, key, make):
        if key in self._d:
            self._d.move_to_end(key)
            return self._d[key]
        while len(self._d) >= self.capacity:
            self._d.popitem(last=False)
        self._d[key] = make()
        return self._d[key]

    def __len__(self):
        return len(self._d)

    def __contains__(self, key):
        return key in self._d


class _Pool:
    """Size-keyed free list so that dead activations are reused (keeps the working set L2-friendly)."""

    def __init__(self, device):
        self.device = device
        self.free_lists: Dict[Tuple[int, torch.dtype], List[torch.Tensor]] = {}
        self.all: List[torch.Tensor] = []       # owns every allocation for the life of the plan
        self.total_bytes = 0
        self.on_recycle = None                  # called with a buffer's address when it is freed or handed out again

    def alloc(self, *shape, dtype=F16) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= s
        fl = self.free_lists.get((n, dtype))
        if fl:
            t = fl.pop()
        else:
            self.total_bytes += n * torch.empty((), dtype=dtype).element_size()
            t = torch.empty(n, dtype=dtype, device=self.device)
            self.all.append(t)
        if self.on_recycle is not None:
            self.on_recycle(t.data_ptr())
        return t.view(*shape)

    def free(self, t: torch.Tensor):
        base = t.reshape(-1)
        if self.on_recycle is not None:
            self.on_recycle(base.data_ptr())
        self.free_lists.setdefault((base.numel(), t.dtype), []).append(base)


class _GnStatsMixin:
    """Bookkeeping for GroupNorm statistics fused into producer epilogues: ``_produce`` records a GEMM (or the four
    phase GEMMs of an upsample conv) and, when its output can feed a two-pass GroupNorm, makes it leave column
    partials in a pooled fp32 buffer that lives exactly as long as the output tensor; ``_colstats`` hands them to the
    consuming GroupNorm.  Needs ``self.pool`` and ``self.prog``."""

    def _init_gn_stats(self):
        self._cs: Dict[int, tuple] = {}          # output address -> (partials [B, slots, C, 2], slots, pooled buffer)
        self.pool.on_recycle = self._drop_stats
        self.tickets = torch.zeros(max(self.B, 1), dtype=torch.int32, device=self.pool.device)

    def _drop_stats(self, ptr: int):
        e = self._cs.pop(ptr, None)
        if e is not None:
            self.pool.free(e[2])

    def _produce(self, descs, out: torch.Tensor):
        """out: the [B, h, w, C] tensor the launch(es) write."""
        Bn, h, w, c = out.shape
        dl = descs if isinstance(descs, (list, tuple)) else [descs]
        if FUSE_GN_STATS and h * w >= GN_STATS_MIN_HW and c % 32 == 0:
            flat = dl[0].H == 1 and dl[0].B == 1 and Bn * h * w == dl[0].W      # a [rows, C] GEMM: slots run over all rows
            per = [K.colstats_slots(d) for d in dl]
            ok = min(per) >= 1 and (not flat or ((h * w) % 32 == 0 and len(dl) == 1))
            if ok:
                total = sum(per)
                raw = self.pool.alloc((1 if flat else Bn) * total * c * 2, dtype=F32)
                base = 0
                for d, n in zip(dl, per):
                    d.colstats_out, d.colstats_slots, d.colstats_slot_base = raw.data_ptr(), total, base
                    base += n
                if flat:
                    slots = (h * w) // 32          # rows are (b, pixel): the first B * HW / 32 slots, b-major
                    view = raw[:Bn * slots * c * 2].view(Bn, slots, c, 2)
                else:
                    slots, view = total, raw.view(Bn, total, c, 2)
                self._cs[out.data_ptr()] = (view, slots, raw)
        self.prog.add(descs)

    def _gn_coef(self, x0, x1, g, b, eps):
        """Statistics-only GroupNorm of [x0 | x1]: per-(sample, channel) (scale, shift) for a halo conv's operand path."""
        c = x0.shape[-1] + (0 if x1 is None else x1.shape[-1])
        cs = self._colstats(x0, x1)
        coef = self.pool.alloc(x0.shape[0], c, 2, dtype=F32)
        self.prog.add(K.group_norm_coef_op(x0, x1, g, b, coef, self.stats, self.tickets, eps, colstats=cs))
        return coef

    @staticmethod
    def _use_halo(x0, x1):
        """The halo-tile conv has no split-K: below HALO_MIN_TILES 128-pixel tiles (small batches on the 16 x 16 / 32 x 32
        levels) the implicit GEMM's K split fills the GPU and wins (batch 1 at 64 x 64: 5.43 vs 4.90 ms per UNet step with
        the halo conv everywhere)."""
        bn, h, w, c0 = x0.shape
        return (HALO and h * w >= HALO_MIN_HW and w % 8 == 0 and h % 16 == 0 and c0 % 8 == 0 and
                bn * (h // 16) * (w // 8) >= HALO_MIN_TILES and
                (x1 is None or (c0 % 64 == 0 and x1.shape[-1] % 8 == 0)))

    @staticmethod
    def _gn_fold_ok(x):
        """Per-sample weights need GEMM tiles (128 pixels, CTA pairs: 256) that stay inside one sample and the two-pass
        GroupNorm regime (small maps take the one-launch gn_small kernel)."""
        bn, h, w, c = x.shape
        cpg = c // 32
        small = cpg % 8 == 0 and h * w * (cpg // 8) <= 2048
        return (not small) and (h * w) % 256 == 0 and w in (8, 16, 32, 64, 128) and c % 64 == 0 and c <= 2048

    def _colstats(self, x0, x1):
        a = self._cs.get(x0.data_ptr())
        b = None if x1 is None else self._cs.get(x1.data_ptr())
        if a is None or (x1 is not None and b is None):
            return None
        return (a[0], a[1]), (None if b is None else (b[0], b[1]))


def _f32(sd, name, dev):
    """fp16-rounded parameter as fp32 (the reference fp16 model stores every parameter in fp16)."""
    return sd[name].detach().to(F16).float().contiguous().to(dev)      # converted where it lives, then one copy


# =============================================================================== UNet
class UNetEngine:
    """SD1.x UNet with reference parameter names.  ``denoise`` has BaseModel.apply_model semantics."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], cfg: UNetConfig = SD15_UNET, device="cuda"):
        L.require_device()
        self.cfg, self.dev = cfg, torch.device(device)
        self.layout = unet_layout(cfg)
        self.sigmas, log_sigmas = sigma_tables()
        self.log_sigmas = log_sigmas.to(self.dev)
        self.w: Dict[str, torch.Tensor] = {}
        self._plans = _PlanCache()
        self._pack(state_dict)

    # ---- one-time weight packing
    def _pack(self, sd):
        dev, w = self.dev, self.w
        w["te0.w"], w["te0.b"] = K.pack_linear(sd["time_embed.0.weight"]), _f32(sd, "time_embed.0.bias", dev)
        w["te2.w"], w["te2.b"] = K.pack_linear(sd["time_embed.2.weight"]), _f32(sd, "time_embed.2.bias", dev)
        emb_w, emb_b, self.emb_off, off = [], [], {}, 0
        kv_w, self.kv_off, kvo = [], {}, 0
        blocks = self.layout.input_blocks + [self.layout.middle] + self.layout.output_blocks
        for block in blocks:
            for s in block:
                if isinstance(s, ResSpec):
                    self._pack_res(sd, s)
                    emb_w.append(sd[s.name + ".emb_layers.1.weight"])
                    emb_b.append(sd[s.name + ".emb_layers.1.bias"])
                    self.emb_off[s.name] = off
                    off += s.cout
                elif isinstance(s, AttnSpec):
                    self._pack_attn(sd, s)
                    t = s.name + ".transformer_blocks.0.attn2"
                    kv_w += [sd[t + ".to_k.weight"], sd[t + ".to_v.weight"]]
                    self.kv_off[s.name] = kvo
                    kvo += 2 * s.ch
                elif s.kind == "up":
                    w[s.name + ".w"] = K.pack_upconv3x3(sd[s.name + ".weight"])
                    w[s.name + ".b"] = _f32(sd, s.name + ".bias", dev)
                else:
                    w[s.name + ".w"] = K.pack_conv3x3(self._pad_in(sd[s.name + ".weight"]))
                    w[s.name + ".b"] = _f32(sd, s.name + ".bias", dev)
        w["emb_all.w"] = K.pack_linear(torch.cat(emb_w, 0))
        w["emb_all.b"] = torch.cat([t.detach().to(F16).float() for t in emb_b]).contiguous().to(dev)
        self.emb_total = off
        w["kv_all.w"] = K.pack_linear(torch.cat(kv_w, 0))
        self.kv_total = kvo
        w["out.g"], w["out.beta"] = _f32(sd, "out.0.weight", dev), _f32(sd, "out.0.bias", dev)
        w["out.w"], w["out.b"] = K.pack_conv3x3(sd["out.2.weight"]), _f32(sd, "out.2.bias", dev)
        if CONV_SMALL and sd["out.2.weight"].shape[0] <= 5:
            w["out.ws"] = K.pack_conv3x3_small(sd["out.2.weight"])
        torch.cuda.synchronize()

    @staticmethod
    def _pad_in(wt):
        """conv_in: pad Cin 4 -> 8 so that the NHWC pixel pitch is 16 bytes (TMA stride rule)."""
        if wt.shape[1] % 8 == 0:
            return wt
        pad = 8 - wt.shape[1] % 8
        return torch.cat([wt, wt.new_zeros(wt.shape[0], pad, 3, 3)], 1)

    def _pack_res(self, sd, s: ResSpec):
        dev, w, p = self.dev, self.w, s.name
        w[p + ".gn1.g"], w[p + ".gn1.b"] = _f32(sd, p + ".in_layers.0.weight", dev), _f32(sd, p + ".in_layers.0.bias", dev)
        w[p + ".conv1.w"] = K.pack_conv3x3(sd[p + ".in_layers.2.weight"])
        w[p + ".conv1.b"] = _f32(sd, p + ".in_layers.2.bias", dev)
        w[p + ".gn2.g"], w[p + ".gn2.b"] = _f32(sd, p + ".out_layers.0.weight", dev), _f32(sd, p + ".out_layers.0.bias", dev)
        w2 = K.pack_conv3x3(sd[p + ".out_layers.3.weight"])
        b2 = _f32(sd, p + ".out_layers.3.bias", dev)
        if s.cin != s.cout:
            # skip_connection (1x1 conv, LD.py:5267) rides along as extra K segments of the second 3x3 conv
            w2 = torch.cat([w2, K.pack_linear(sd[p + ".skip_connection.weight"])], 1).contiguous()
            b2 = b2 + _f32(sd, p + ".skip_connection.bias", dev)
        w[p + ".conv2.w"], w[p + ".conv2.b"] = w2, b2

    def _pack_attn(self, sd, s: AttnSpec):
        dev, w, p = self.dev, self.w, s.name
        t = p + ".transformer_blocks.0"
        w[p + ".gn.g"], w[p + ".gn.b"] = _f32(sd, p + ".norm.weight", dev), _f32(sd, p + ".norm.bias", dev)
        w[p + ".pin.w"], w[p + ".pin.b"] = K.pack_linear(sd[p + ".proj_in.weight"]), _f32(sd, p + ".proj_in.bias", dev)
        w[p + ".pout.w"], w[p + ".pout.b"] = K.pack_linear(sd[p + ".proj_out.weight"]), _f32(sd, p + ".proj_out.bias", dev)
        # The three LayerNorms of BasicTransformerBlock (LD.py:4117-4162) are folded into the GEMMs that consume
        # them (LDX_EPI_ROWNORM): w' = w * gamma, gsum = rowsum(w'), bias' = bias + w @ beta.
        ln = lambda n: (sd[f"{t}.{n}.weight"].to(F16).float(), sd[f"{t}.{n}.bias"].to(F16).float())
        wqkv = torch.cat([sd[t + ".attn1.to_q.weight"], sd[t + ".attn1.to_k.weight"], sd[t + ".attn1.to_v.weight"]], 0)
        w[p + ".qkv.w"], w[p + ".qkv.g"], w[p + ".qkv.b"] = K.fold_layer_norm(wqkv, None, *ln("norm1"))
        w[p + ".o1.w"], w[p + ".o1.b"] = K.pack_linear(sd[t + ".attn1.to_out.0.weight"]), _f32(sd, t + ".attn1.to_out.0.bias", dev)
        w[p + ".q2.w"], w[p + ".q2.g"], w[p + ".q2.b"] = K.fold_layer_norm(sd[t + ".attn2.to_q.weight"], None, *ln("norm2"))
        w[p + ".o2.w"], w[p + ".o2.b"] = K.pack_linear(sd[t + ".attn2.to_out.0.weight"]), _f32(sd, t + ".attn2.to_out.0.bias", dev)
        if FUSE_LN3:
            wf, gf, bf = K.fold_layer_norm(sd[t + ".ff.net.0.proj.weight"], sd[t + ".ff.net.0.proj.bias"].to(F16).float(),
                                           *ln("norm3"))
            w[p + ".ff1.w"], w[p + ".ff1.b"] = K.pack_geglu(wf, bf, 128)
            _, w[p + ".ff1.g"] = K.pack_geglu(wf, gf, 128)
            if FF_FUSED and s.ch == 320:           # the fused feed-forward kernel interleaves value / gate rows by 64
                w[p + ".ff1.w64"], w[p + ".ff1.b64"] = K.pack_geglu(wf, bf, 64)
                _, w[p + ".ff1.g64"] = K.pack_geglu(wf, gf, 64)
        else:
            w[p + ".norm3.g"], w[p + ".norm3.b"] = _f32(sd, t + ".norm3.weight", dev), _f32(sd, t + ".norm3.bias", dev)
            w[p + ".ff1.w"], w[p + ".ff1.b"] = K.pack_geglu(sd[t + ".ff.net.0.proj.weight"],
                                                           sd[t + ".ff.net.0.proj.bias"].to(F16).float(), 128)
        w[p + ".ff2.w"], w[p + ".ff2.b"] = K.pack_linear(sd[t + ".ff.net.2.weight"]), _f32(sd, t + ".ff.net.2.bias", dev)

    # ---- plans
    def plan(self, B: int, H: int, W: int, cfg_pairs: bool = False, ctx_len: int = 77) -> "UNetPlan":
        """ctx_len: context tokens per sample (77 * sections: the reference concatenates the 77-token sections of a
        long prompt and repeats the shorter of cond / uncond to their lcm, LD.py:647-663)."""
        return self._plans.get((B, H, W, cfg_pairs, ctx_len), lambda: UNetPlan(self, B, H, W, cfg_pairs, ctx_len))

    @staticmethod
    def cat_contexts(*ctxs: torch.Tensor) -> torch.Tensor:
        """CONDCrossAttn.concat (LD.py:647-663): contexts of different token counts are repeated to their lcm
        (repeating keys does not change softmax attention), then concatenated along the batch."""
        n = 1
        for c in ctxs:
            n = n * c.shape[1] // math.gcd(n, c.shape[1])
        return torch.cat([c if c.shape[1] == n else c.repeat(1, n // c.shape[1], 1) for c in ctxs])

    @torch.no_grad()
    def denoise(self, x: torch.Tensor, sigma: torch.Tensor, context: Optional[torch.Tensor]) -> torch.Tensor:
        """apply_model: x fp32 [B,4,H,W], sigma [B], context [B,77,ctx] -> denoised fp32 [B,4,H,W]."""
        B, _, H, W = x.shape
        return self.plan(B, H, W, ctx_len=context.shape[1]).run(x, sigma, context).clone()

    @torch.no_grad()
    def eps(self, x, sigma, context) -> torch.Tensor:
        """Raw UNet output (fp16 NHWC buffer converted to fp32 NCHW) for parity tests."""
        B, _, H, W = x.shape
        pl = self.plan(B, H, W, ctx_len=context.shape[1])
        pl.run(x, sigma, context)
        return pl.eps_buf[..., :self.cfg.out_channels].float().permute(0, 3, 1, 2).contiguous()

    @torch.no_grad()
    def denoise_cfg(self, x, sigma, cond, uncond, scale: float, context_changed: bool = True) -> torch.Tensor:
        """sampling_function (LD.py:2609): one 2b batch ordered [uncond, cond], then cfg_function."""
        b, _, H, W = x.shape
        ctx = self.cat_contexts(uncond, cond)
        pl = self.plan(2 * b, H, W, cfg_pairs=True, ctx_len=ctx.shape[1])
        return pl.run(x, sigma, ctx if context_changed else None, cfg_scale=scale).clone()


class UNetPlan(_GnStatsMixin):
    def __init__(self, eng: UNetEngine, B: int, H: int, W: int, cfg_pairs: bool, ctx_len: int = 77):
        self.eng, self.B, self.H, self.W, self.cfg_pairs, self.ctx_len = eng, B, H, W, cfg_pairs, ctx_len
        if ctx_len < 1:
            raise L.LdxError(f"context length {ctx_len} must be positive")
        cfg, dev = eng.cfg, eng.dev
        down = 2 ** (len(cfg.channel_mult) - 1)
        if H % down or W % down:
            raise L.LdxError(f"latent {H}x{W} must be a multiple of {down}")
        self.pool = _Pool(dev)
        self._init_gn_stats()
        nb = B // 2 if cfg_pairs else B          # number of distinct latents fed in
        self.nb = nb
        self.x_in = torch.zeros(nb, cfg.in_channels, H, W, dtype=F32, device=dev)
        self.sigma_in = torch.ones(nb, dtype=F32, device=dev)
        self.ctx = torch.zeros(B, ctx_len, cfg.context_dim, dtype=F16, device=dev)
        self.den = torch.zeros(nb, cfg.out_channels, H, W, dtype=F32, device=dev)
        self.cfg_scale = torch.tensor([1.0])    # host copy; the CFG op is re-recorded when it changes
        self.stats = torch.empty(B * L.GN_STATS_FLOATS_PER_BATCH, dtype=F32, device=dev)
        self._rs_cache = {}
        self.kv_all = torch.empty(B * ctx_len, eng.kv_total, dtype=F16, device=dev)
        self.ctx_prog = K.Program()
        self.ctx_prog.add(K.linear_desc(self.ctx.view(B * ctx_len, -1), eng.w["kv_all.w"], self.kv_all))
        self.prog = K.Program()
        self._build()
        self.tail_prog: Optional[K.Program] = None
        self._tail_scale: Optional[float] = None
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.use_graph = True

    # ---- op helpers
    def _gn(self, x0, x1, g, b, eps, silu):
        Bn, h, w, c0 = x0.shape
        c = c0 + (0 if x1 is None else x1.shape[-1])
        cs = self._colstats(x0, x1)               # looked up before y is allocated (allocation recycles buffers)
        y = self.pool.alloc(Bn, h, w, c)
        self.prog.add(K.group_norm_op(x0, x1, g, b, y, self.stats, eps, silu, colstats=cs))
        return y

    def _res(self, s: ResSpec, x0, x1):
        eng, w, p = self.eng, self.eng.w, s.name
        Bn, h, wd, _ = x0.shape
        if self._use_halo(x0, x1):
            return self._res_halo(s, x0, x1)
        y = self._gn(x0, x1, w[p + ".gn1.g"], w[p + ".gn1.b"], 1e-5, True)
        h1 = self.pool.alloc(Bn, h, wd, s.cout)
        off = eng.emb_off[p]
        self._produce(K.conv3x3_desc(y, w[p + ".conv1.w"], h1, bias=w[p + ".conv1.b"],
                                     rowadd=self.emb_out[:, off:off + s.cout]), h1)
        self.pool.free(y)
        y2 = self._gn(h1, None, w[p + ".gn2.g"], w[p + ".gn2.b"], 1e-5, True)
        self.pool.free(h1)
        out = self.pool.alloc(Bn, h, wd, s.cout)
        nch = K.ceil64(s.cout) // 64
        segs = [(0, 0, kx - 1, ky - 1, nch) for ky in range(3) for kx in range(3)]
        if s.cin == s.cout:
            assert x1 is None
            d = K.gemm_desc([y2], segs, w[p + ".conv2.w"], out, s.cout, bias=w[p + ".conv2.b"], residual=x0)
        else:
            srcs = [y2, x0] + ([x1] if x1 is not None else [])
            c0 = x0.shape[-1]
            assert x1 is None or c0 % 64 == 0
            segs.append((1, 0, 0, 0, K.ceil64(c0) // 64))
            if x1 is not None:
                segs.append((2, 0, 0, 0, K.ceil64(x1.shape[-1]) // 64))
            d = K.gemm_desc(srcs, segs, w[p + ".conv2.w"], out, s.cout, bias=w[p + ".conv2.b"])
        self._produce(d, out)
        self.pool.free(y2)
        return out

    def _res_halo(self, s: ResSpec, x0, x1):
        """ResBlock1 with both GroupNorm + SiLU applications inside the convs' operand paths: two statistics launches
        and two halo-tile convs; the normalised maps never exist in HBM."""
        eng, w, p, pool = self.eng, self.eng.w, s.name, self.pool
        Bn, h, wd, c0 = x0.shape
        xs = [x0] + ([x1] if x1 is not None else [])
        coef1 = self._gn_coef(x0, x1, w[p + ".gn1.g"], w[p + ".gn1.b"], 1e-5)
        h1 = pool.alloc(Bn, h, wd, s.cout)
        off = eng.emb_off[p]
        self._produce(K.conv3x3_halo_desc(xs, w[p + ".conv1.w"], h1, gn=(coef1, True), bias=w[p + ".conv1.b"],
                                          rowadd=self.emb_out[:, off:off + s.cout]), h1)
        pool.free(coef1)
        coef2 = self._gn_coef(h1, None, w[p + ".gn2.g"], w[p + ".gn2.b"], 1e-5)
        out = pool.alloc(Bn, h, wd, s.cout)
        if s.cin == s.cout:
            assert x1 is None
            d = K.conv3x3_halo_desc([h1], w[p + ".conv2.w"], out, gn=(coef2, True), bias=w[p + ".conv2.b"], residual=x0)
        else:
            segs = [(i, 0, 0, 0, K.ceil64(t.shape[-1]) // 64) for i, t in enumerate(xs)]
            d = K.conv3x3_halo_desc([h1], w[p + ".conv2.w"], out, gn=(coef2, True), bias=w[p + ".conv2.b"],
                                    extra_srcs=xs, extra_segs=segs)
        self._produce(d, out)
        pool.free(coef2); pool.free(h1)
        return out

    def _rowstats(self, d, rows: int, slot: int):
        """[rows, parts, 2] fp32 buffer (cached per shape, two alternating slots) that GEMM `d` fills with the
        per-row (sum, sum of squares) of its outputs for the LayerNorm folded into the next GEMM."""
        parts = K.rowstats_parts(d)
        key = (rows, parts, slot)
        if key not in self._rs_cache:
            self._rs_cache[key] = torch.empty(rows, parts, 2, dtype=F32, device=self.eng.dev)
        st = self._rs_cache[key]
        d.rowstats_out, d.rowstats_out_parts = st.data_ptr(), parts
        return st

    def _rowstats_like(self, st: torch.Tensor, rows: int, slot: int):
        key = (rows, st.shape[1], slot, "dup")
        if key not in self._rs_cache:
            self._rs_cache[key] = torch.empty(rows, st.shape[1], 2, dtype=F32, device=self.eng.dev)
        return self._rs_cache[key]

    def _dup(self, src: torch.Tensor, dst: torch.Tensor):
        """dst = cat([src, src]) along the batch axis (two device copies recorded in the program)."""
        assert src.is_contiguous() and dst.is_contiguous() and dst.numel() == 2 * src.numel() and dst.dtype == src.dtype
        self.prog.add(K.generic_op(L.OP_DUP, p=[src.data_ptr(), dst.data_ptr()], l=[src.numel() * src.element_size()]))

    def _attn(self, s: AttnSpec, x, expand: bool = False):
        """SpatialTransformer (LD.py:4242-4262).  expand=True: x holds only the b distinct latents of a CFG pair
        batch; everything up to and including the self-attention residual is identical for the uncond and cond
        halves (same x, same timestep; only the context differs), so it runs once on b samples and the
        activations are duplicated to the [uncond, cond] batch right before the first cross-attention."""
        eng, w, p, pool = self.eng, self.eng.w, s.name, self.pool
        Bn, h, wd, c = x.shape
        n, m = h * wd, Bn * h * wd
        t = pool.alloc(m, c)
        if HALO_1X1 and self._use_halo(x, None):
            # SpatialTransformer.norm folded into proj_in's operand path (statistics-only GroupNorm launch)
            coef = self._gn_coef(x, None, w[p + ".gn.g"], w[p + ".gn.b"], 1e-6)
            d = K.conv1x1_gn_desc(x, w[p + ".pin.w"], t.view(Bn, h, wd, c), gn=(coef, False), bias=w[p + ".pin.b"])
            st1 = self._rowstats(d, m, 0)      # row statistics of t for norm1, folded into the QKV GEMM
            self.prog.add(d)
            pool.free(coef)
        elif GN_FOLD and self._gn_fold_ok(x):
            # SpatialTransformer.norm folded into proj_in's WEIGHTS: statistics-only GroupNorm launch, one small launch that
            # scales proj_in's weight per sample (w'[b] = w * scale[b], bias'[b]), then the 1x1 conv reads x itself: the
            # normalised map is never written or read (gn_apply: 2 x the map per block)
            coef = self._gn_coef(x, None, w[p + ".gn.g"], w[p + ".gn.b"], 1e-6)
            wb = pool.alloc(Bn * c, w[p + ".pin.w"].stride(0)).view(Bn, c, w[p + ".pin.w"].stride(0))
            bb = torch.empty(Bn, c, dtype=torch.float32, device=x.device)
            self.prog.add(K.gn_fold_weights_op(w[p + ".pin.w"], w[p + ".pin.b"], w[p + ".gn.b"], coef, wb, bb), keep=[bb])
            d = K.conv1x1_batched_desc(x, wb, bb, t.view(Bn, h, wd, c))
            st1 = self._rowstats(d, m, 0)      # row statistics of t for norm1, folded into the QKV GEMM
            self.prog.add(d)
            pool.free(coef)
            pool.free(wb)
        else:
            y = self._gn(x, None, w[p + ".gn.g"], w[p + ".gn.b"], 1e-6, False)
            d = K.linear_desc(y, w[p + ".pin.w"], t, bias=w[p + ".pin.b"])
            st1 = self._rowstats(d, m, 0)      # row statistics of t for norm1, folded into the QKV GEMM
            self.prog.add(d)
            pool.free(y)
        # attn1 (self)
        qkv = pool.alloc(m, 3 * c)
        self.prog.add(K.linear_desc(t, w[p + ".qkv.w"], qkv, bias=w[p + ".qkv.b"], rownorm=(st1, w[p + ".qkv.g"], 1e-5, c)))
        a = pool.alloc(m, c)
        q3 = qkv.view(Bn, n, 3 * c)
        self.prog.add(K.attention_op(q3[:, :, :c], q3[:, :, c:2 * c], q3[:, :, 2 * c:], a.view(Bn, n, c), s.heads))
        pool.free(qkv)
        t2 = pool.alloc(m, c)
        d = K.linear_desc(a, w[p + ".o1.w"], t2, bias=w[p + ".o1.b"], residual=t)
        st2 = self._rowstats(d, m, 1)          # norm2 -> attn2.to_q
        self.prog.add(d)
        pool.free(t)
        if expand:                             # b -> [uncond, cond] = 2b samples from here on
            pool.free(a)
            t2b, xb = pool.alloc(2 * m, c), pool.alloc(2 * Bn, h, wd, c)
            st2b = self._rowstats_like(st2, 2 * m, 1)
            self._dup(t2, t2b)
            self._dup(st2, st2b)
            self._dup(x, xb)
            pool.free(t2)
            t2, st2, x, Bn, m = t2b, st2b, xb, 2 * Bn, 2 * m
            a = pool.alloc(m, c)
        # attn2 (cross; K/V of the step-invariant context come from the kv_all GEMM)
        q2 = a
        self.prog.add(K.linear_desc(t2, w[p + ".q2.w"], q2, bias=w[p + ".q2.b"], rownorm=(st2, w[p + ".q2.g"], 1e-5, c)))
        kvo = eng.kv_off[p]
        kv3 = self.kv_all.view(Bn, self.ctx_len, eng.kv_total)
        a2 = pool.alloc(m, c)
        self.prog.add(K.attention_op(q2.view(Bn, n, c), kv3[:, :, kvo:kvo + c], kv3[:, :, kvo + c:kvo + 2 * c],
                                     a2.view(Bn, n, c), s.heads))
        t3 = pool.alloc(m, c)
        d = K.linear_desc(a2, w[p + ".o2.w"], t3, bias=w[p + ".o2.b"], residual=t2)
        st3 = self._rowstats(d, m, 0) if FUSE_LN3 else None     # norm3 -> GEGLU projection
        self.prog.add(d)
        pool.free(a2); pool.free(t2); pool.free(q2)
        # feed-forward (the GEGLU gate lives in the first GEMM's epilogue)
        if (p + ".ff1.w64") in w:
            t4 = pool.alloc(m, c)
            self.prog.add(K.ff_fused_op(t3, w[p + ".ff1.w64"], w[p + ".ff1.b64"], w[p + ".ff1.g64"], w[p + ".ff2.w"],
                                        w[p + ".ff2.b"], st3, t4))
            pool.free(t3)
            out = pool.alloc(Bn, h, wd, c)
            self._produce(K.linear_desc(t4, w[p + ".pout.w"], out, bias=w[p + ".pout.b"], residual=x), out)
            pool.free(t4)
            if expand:
                pool.free(x)
            return out
        g = pool.alloc(m, 4 * c)
        if FUSE_LN3:
            self.prog.add(K.linear_desc(t3, w[p + ".ff1.w"], g, bias=w[p + ".ff1.b"], flags=L.EPI_GEGLU, bn=256,
                                        rownorm=(st3, w[p + ".ff1.g"], 1e-5, c)))
        else:
            nrm = pool.alloc(m, c)
            self.prog.add(K.layer_norm_op(t3, w[p + ".norm3.g"], w[p + ".norm3.b"], nrm))
            self.prog.add(K.linear_desc(nrm, w[p + ".ff1.w"], g, bias=w[p + ".ff1.b"], flags=L.EPI_GEGLU, bn=256))
            pool.free(nrm)
        t4 = pool.alloc(m, c)
        self.prog.add(K.linear_desc(g, w[p + ".ff2.w"], t4, bias=w[p + ".ff2.b"], residual=t3))
        pool.free(g); pool.free(t3)
        out = pool.alloc(Bn, h, wd, c)
        self._produce(K.linear_desc(t4, w[p + ".pout.w"], out, bias=w[p + ".pout.b"], residual=x), out)
        pool.free(t4)
        if expand:
            pool.free(x)                       # the duplicated copy made above
        return out

    def _block(self, layers, h, hskip=None):
        """Runs one TimestepEmbedSequential (LD.py:5089-5111). Returns the new h (inputs are not freed)."""
        pool = self.pool
        cur, owned = h, False
        for s in layers:
            if isinstance(s, ResSpec):
                nxt = self._res(s, cur, hskip)
                hskip = None
            elif isinstance(s, AttnSpec):
                nxt = self._attn(s, cur)
            elif s.kind == "conv_in":
                Bn, hh, ww, _ = cur.shape
                nxt = pool.alloc(Bn, hh, ww, s.cout)
                self._produce(K.conv3x3_desc(cur, self.eng.w[s.name + ".w"], nxt, bias=self.eng.w[s.name + ".b"]), nxt)
            elif s.kind == "down":
                Bn, hh, ww, _ = cur.shape
                nxt = pool.alloc(Bn, hh // 2, ww // 2, s.cout)
                self._produce(K.conv3x3_s2_desc(cur, self.eng.w[s.name + ".w"], nxt, bias=self.eng.w[s.name + ".b"]), nxt)
            else:  # up: nearest 2x + conv folded into four phase convs
                Bn, hh, ww, _ = cur.shape
                nxt = pool.alloc(Bn, hh * 2, ww * 2, s.cout)
                self._produce(K.upconv3x3_descs(cur, self.eng.w[s.name + ".w"], nxt, bias=self.eng.w[s.name + ".b"]), nxt)
            if owned:
                pool.free(cur)
            cur, owned = nxt, True
        return cur

    def _build(self):
        eng, cfg, w, pool, B, H, W = self.eng, self.eng.cfg, self.eng.w, self.pool, self.B, self.H, self.W
        prog = self.prog
        cpad, nb = 8, self.nb
        x0 = torch.zeros(B, H, W, cpad, dtype=F16, device=eng.dev)
        temb = torch.empty(B, cfg.model_channels, dtype=F16, device=eng.dev)
        # with CFG the same latents feed both halves of the [uncond, cond] batch
        # (torch.cat([x] * 2), torch.cat([timestep] * 2) in calc_cond_batch, LD.py:2545-2551)
        for half in range(B // nb):
            prog.add(K.generic_op(L.OP_NCHW_TO_NHWC,
                                  p=[self.x_in.data_ptr(), self.sigma_in.data_ptr(), x0[half * nb:].data_ptr()],
                                  i=[nb, cfg.in_channels, H, W, cpad]))
            prog.add(K.generic_op(L.OP_TIMESTEP_EMB,
                                  p=[self.sigma_in.data_ptr(), eng.log_sigmas.data_ptr(), temb[half * nb:].data_ptr()],
                                  i=[1000, nb, cfg.model_channels]))
        e1 = torch.empty(B, cfg.time_embed_dim, dtype=F16, device=eng.dev)
        semb = torch.empty(B, cfg.time_embed_dim, dtype=F16, device=eng.dev)
        prog.add(K.linear_desc(temb, w["te0.w"], e1, bias=w["te0.b"], flags=L.EPI_SILU))
        # emb is only ever consumed through SiLU (emb_layers = [SiLU, Linear], LD.py:5237-5248)
        prog.add(K.linear_desc(e1, w["te2.w"], semb, bias=w["te2.b"], flags=L.EPI_SILU))
        self.emb_out = torch.empty(B, eng.emb_total, dtype=F32, device=eng.dev)
        prog.add(K.linear_desc(semb, w["emb_all.w"], self.emb_out, bias=w["emb_all.b"], flags=L.EPI_OUT_F32))
        prog.keep(x0, temb, e1, semb)

        hs = []
        h = x0
        blocks = list(eng.layout.input_blocks)
        l1 = blocks[1] if len(blocks) > 1 else []
        share = (self.cfg_pairs and SHARE_CFG_PREFIX and len(blocks) > 1 and len(l1) == 2 and
                 isinstance(l1[0], ResSpec) and isinstance(l1[1], AttnSpec))
        if share:
            # CFG: conv_in, the first ResBlock and the first self-attention see identical inputs in both halves
            h0 = self._block(blocks[0], x0[:nb])                 # conv_in on the b distinct latents
            r1 = self._res(l1[0], h0, None)
            h = self._attn(l1[1], r1, expand=True)
            pool.free(r1)
            h0b = pool.alloc(B, H, W, h0.shape[-1])
            self._dup(h0, h0b)
            e = self._cs.get(h0.data_ptr())
            if e is not None:                                   # the statistics of the duplicated batch are duplicates too
                raw = pool.alloc(2 * e[0].numel(), dtype=F32)
                self._dup(e[0], raw)
                self._cs[h0b.data_ptr()] = (raw.view(B, e[1], h0.shape[-1], 2), e[1], raw)
            pool.free(h0)
            hs += [h0b, h]
            blocks = blocks[2:]
        for layers in blocks:
            h = self._block(layers, h)
            hs.append(h)
        h = self._block(eng.layout.middle, h)
        for layers in eng.layout.output_blocks:
            hsp = hs.pop()
            nh = self._block(layers, h, hsp)
            if not any(h is t for t in hs):
                pool.free(h)
            pool.free(hsp)
            h = nh
        y = self._gn(h, None, w["out.g"], w["out.beta"], 1e-5, True)
        self.eps_buf = torch.zeros(B, H, W, 8, dtype=F16, device=eng.dev)
        if "out.ws" in w:
            prog.add(K.conv3x3_small_op(y, w["out.ws"], w["out.b"], cfg.out_channels, out16=self.eps_buf))
        else:
            prog.add(K.conv3x3_desc(y, w["out.w"], self.eps_buf[..., :cfg.out_channels], bias=w["out.b"]))
        if not self.cfg_pairs:
            prog.add(K.generic_op(L.OP_EPS_TO_DENOISED,
                                  p=[self.eps_buf.data_ptr(), self.x_in.data_ptr(), self.sigma_in.data_ptr(),
                                     self.den.data_ptr()],
                                  i=[B, cfg.out_channels, H, W], l=[8]))
        self.act_bytes = pool.total_bytes

    def _tail(self, scale: float) -> K.Program:
        if self.tail_prog is None or self._tail_scale != scale:
            cfg = self.eng.cfg
            tp = K.Program()
            tp.add(K.generic_op(L.OP_CFG_DENOISED,
                                p=[self.eps_buf.data_ptr(), self.x_in.data_ptr(), self.sigma_in.data_ptr(),
                                   self.den.data_ptr()],
                                i=[self.nb, cfg.out_channels, self.H, self.W], l=[8], f=[scale]))
            self.tail_prog, self._tail_scale = tp, scale
        return self.tail_prog

    @property
    def num_launches(self) -> int:
        return self.prog.num_launches + (1 if self.cfg_pairs else 0)

    def set_context(self, context: torch.Tensor):
        """Copies the context in and recomputes the cross-attention K/V of all 16 blocks (step-invariant)."""
        if tuple(context.shape) != tuple(self.ctx.shape):
            raise L.LdxError(f"context must be {tuple(self.ctx.shape)}, got {tuple(context.shape)}")
        self.ctx.copy_(context.to(F16))
        self.ctx_prog.run()

    def run_body(self):
        """Replays the recorded UNet program (CUDA graph) on the plan's resident inputs (x_in, sigma_in, context);
        the result is the fp16 NHWC eps buffer.  The CFG / sampler tail is launched by the caller."""
        if not self.use_graph:
            self.prog.run()
            return
        if self.graph is None:
            self.prog.run()                         # warm-up (sets kernel attributes) outside capture
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.prog.run()
            self.graph = g
        self.graph.replay()

    def run(self, x, sigma, context=None, cfg_scale: Optional[float] = None) -> torch.Tensor:
        """Returns the plan's output buffer (denoised fp32 NCHW); valid until the next run."""
        if context is not None:
            self.set_context(context)
        if x.data_ptr() != self.x_in.data_ptr():
            self.x_in.copy_(x)
        if sigma.data_ptr() != self.sigma_in.data_ptr():
            self.sigma_in.copy_(sigma)
        self.run_body()
        if self.cfg_pairs:
            self._tail(1.0 if cfg_scale is None else float(cfg_scale)).run()
        return self.den


# =============================================================================== VAE decoder
class VAEDecoderEngine:
    """AutoencoderKL decoder (post_quant_conv + Decoder) with reference parameter names."""

    def __init__(self, state_dict, cfg: VAEConfig = SD15_VAE, device="cuda"):
        L.require_device()
        self.cfg, self.dev = cfg, torch.device(device)
        self.w: Dict[str, torch.Tensor] = {}
        self._plans = _PlanCache()
        sd, dev, w = state_dict, self.dev, self.w
        w["pq.w"], w["pq.b"] = K.pack_linear(sd["post_quant_conv.weight"]), _f32(sd, "post_quant_conv.bias", dev)
        w["cin.w"] = K.pack_conv3x3(UNetEngine._pad_in(sd["decoder.conv_in.weight"]))
        w["cin.b"] = _f32(sd, "decoder.conv_in.bias", dev)
        self._pack_res(sd, "decoder.mid.block_1")
        self._pack_res(sd, "decoder.mid.block_2")
        a = "decoder.mid.attn_1"
        w[a + ".gn.g"], w[a + ".gn.b"] = _f32(sd, a + ".norm.weight", dev), _f32(sd, a + ".norm.bias", dev)
        w[a + ".qk.w"] = K.pack_linear(torch.cat([sd[a + ".q.weight"], sd[a + ".k.weight"]], 0))
        w[a + ".qk.b"] = torch.cat([_f32(sd, a + ".q.bias", dev), _f32(sd, a + ".k.bias", dev)]).contiguous()
        w[a + ".v.w"], w[a + ".v.b"] = K.pack_linear(sd[a + ".v.weight"]), _f32(sd, a + ".v.bias", dev)
        w[a + ".o.w"], w[a + ".o.b"] = K.pack_linear(sd[a + ".proj_out.weight"]), _f32(sd, a + ".proj_out.bias", dev)
        for i_level, blocks, has_up in vae_decoder_levels(cfg):
            for j in range(len(blocks)):
                self._pack_res(sd, f"decoder.up.{i_level}.block.{j}")
            if has_up:
                n = f"decoder.up.{i_level}.upsample.conv"
                w[n + ".w"], w[n + ".b"] = K.pack_upconv3x3(sd[n + ".weight"]), _f32(sd, n + ".bias", dev)
        w["nout.g"], w["nout.b"] = _f32(sd, "decoder.norm_out.weight", dev), _f32(sd, "decoder.norm_out.bias", dev)
        w["cout.w"], w["cout.b"] = K.pack_conv3x3(sd["decoder.conv_out.weight"]), _f32(sd, "decoder.conv_out.bias", dev)
        if CONV_SMALL and sd["decoder.conv_out.weight"].shape[0] <= 5:
            w["cout.ws"] = K.pack_conv3x3_small(sd["decoder.conv_out.weight"])
        torch.cuda.synchronize()

    def _pack_res(self, sd, p):
        dev, w = self.dev, self.w
        w[p + ".gn1.g"], w[p + ".gn1.b"] = _f32(sd, p + ".norm1.weight", dev), _f32(sd, p + ".norm1.bias", dev)
        w[p + ".conv1.w"], w[p + ".conv1.b"] = K.pack_conv3x3(sd[p + ".conv1.weight"]), _f32(sd, p + ".conv1.bias", dev)
        w[p + ".gn2.g"], w[p + ".gn2.b"] = _f32(sd, p + ".norm2.weight", dev), _f32(sd, p + ".norm2.bias", dev)
        w2, b2 = K.pack_conv3x3(sd[p + ".conv2.weight"]), _f32(sd, p + ".conv2.bias", dev)
        if (p + ".nin_shortcut.weight") in sd:
            w2 = torch.cat([w2, K.pack_linear(sd[p + ".nin_shortcut.weight"])], 1).contiguous()
            b2 = b2 + _f32(sd, p + ".nin_shortcut.bias", dev)
        w[p + ".conv2.w"], w[p + ".conv2.b"] = w2, b2

    def plan(self, B, H, W) -> "VAEPlan":
        return self._plans.get((B, H, W), lambda: VAEPlan(self, B, H, W))

    @torch.no_grad()
    def decode(self, z: torch.Tensor) -> torch.Tensor:
        """VAE.decode: z fp32 [B,4,h,w] (already / 0.18215) -> image fp32 [B, 8h, 8w, 3] in [0,1]."""
        B, _, H, W = z.shape
        return self.plan(B, H, W).run(z).clone()

    @torch.no_grad()
    def decode_u8(self, z: torch.Tensor) -> torch.Tensor:
        """VAE.decode followed by the quantisation the reference applies when it saves an image
        (clip(255 * img, 0, 255).astype(uint8), LD.py:6750-6751): uint8 [B, 8h, 8w, 3]."""
        B, _, H, W = z.shape
        pl = self.plan(B, H, W)
        pl.run(z)
        out = torch.empty(pl.img.shape, dtype=torch.uint8, device=self.dev)
        L.check(L.lib().ldx_vae_postprocess_u8(pl.dec_buf.data_ptr(), pl.dec_buf.shape[-1], out.data_ptr(),
                                               out.numel() // 3, torch.cuda.current_stream().cuda_stream),
                "ldx_vae_postprocess_u8")
        return out

    @torch.no_grad()
    def decode_raw(self, z):
        B, _, H, W = z.shape
        pl = self.plan(B, H, W)
        pl.run(z)
        return pl.dec_buf[..., :3].float().permute(0, 3, 1, 2).contiguous()


class VAEPlan(_GnStatsMixin):
    def __init__(self, eng: VAEDecoderEngine, B, H, W):
        self.eng, self.B, self.H, self.W = eng, B, H, W
        dev, cfg, w = eng.dev, eng.cfg, eng.w
        self.pool = pool = _Pool(dev)
        self.prog = prog = K.Program()
        self._init_gn_stats()
        self.z_in = torch.zeros(B, cfg.z_channels, H, W, dtype=F32, device=dev)
        up = 2 ** (len(cfg.ch_mult) - 1)
        self.img = torch.zeros(B, H * up, W * up, 3, dtype=F32, device=dev)
        self.stats = torch.empty(B * L.GN_STATS_FLOATS_PER_BATCH, dtype=F32, device=dev)
        z0 = torch.zeros(B, H, W, 8, dtype=F16, device=dev)
        prog.add(K.generic_op(L.OP_NCHW_TO_NHWC, p=[self.z_in.data_ptr(), None, z0.data_ptr()],
                              i=[B, cfg.z_channels, H, W, 8]))
        z1 = torch.zeros(B, H, W, 8, dtype=F16, device=dev)       # channels 4..7 stay zero
        prog.add(K.linear_desc(z0, w["pq.w"], z1[..., :cfg.z_channels], bias=w["pq.b"]))
        top = cfg.ch * cfg.ch_mult[-1]
        h = pool.alloc(B, H, W, top)
        self._produce(K.conv3x3_desc(z1, w["cin.w"], h, bias=w["cin.b"]), h)
        prog.keep(z0, z1)
        h = self._res("decoder.mid.block_1", h, top, top)
        h = self._attn("decoder.mid.attn_1", h)
        h = self._res("decoder.mid.block_2", h, top, top)
        for i_level, blocks, has_up in vae_decoder_levels(cfg):
            for j, (cin, cout) in enumerate(blocks):
                h = self._res(f"decoder.up.{i_level}.block.{j}", h, cin, cout)
            if has_up:
                n = f"decoder.up.{i_level}.upsample.conv"
                Bn, hh, ww, c = h.shape
                nh = pool.alloc(Bn, hh * 2, ww * 2, c)
                self._produce(K.upconv3x3_descs(h, w[n + ".w"], nh, bias=w[n + ".b"]), nh)
                pool.free(h)
                h = nh
        Bn, hh, ww, c = h.shape
        cs = self._colstats(h, None)
        y = pool.alloc(Bn, hh, ww, c)
        prog.add(K.group_norm_op(h, None, w["nout.g"], w["nout.b"], y, self.stats, 1e-6, True, colstats=cs))
        self.dec_buf = torch.zeros(Bn, hh, ww, 8, dtype=F16, device=dev)
        if "cout.ws" in w and cfg.out_ch == 3:
            # conv_out and VAE.process_output in one launch: the 128-channel map is read once, the fp16 result
            # (decode_raw / decode_u8) and the clamped fp32 image are both written from the epilogue
            prog.add(K.conv3x3_small_op(y, w["cout.ws"], w["cout.b"], cfg.out_ch, out16=self.dec_buf, out_img=self.img,
                                        post=1))
        else:
            prog.add(K.conv3x3_desc(y, w["cout.w"], self.dec_buf[..., :cfg.out_ch], bias=w["cout.b"]))
            prog.add(K.generic_op(L.OP_VAE_POST, p=[self.dec_buf.data_ptr(), self.img.data_ptr()],
                                  l=[8, Bn * hh * ww]))
        self.graph = None
        self.use_graph = True

    def _res(self, p, x, cin, cout):
        w, pool, prog = self.eng.w, self.pool, self.prog
        Bn, h, wd, _ = x.shape
        if self._use_halo(x, None):
            # ResnetBlock (LD.py:3562-3576) with the two GroupNorm + SiLU applications in the convs' operand paths
            coef1 = self._gn_coef(x, None, w[p + ".gn1.g"], w[p + ".gn1.b"], 1e-6)
            h1 = pool.alloc(Bn, h, wd, cout)
            self._produce(K.conv3x3_halo_desc([x], w[p + ".conv1.w"], h1, gn=(coef1, True), bias=w[p + ".conv1.b"]), h1)
            pool.free(coef1)
            coef2 = self._gn_coef(h1, None, w[p + ".gn2.g"], w[p + ".gn2.b"], 1e-6)
            out = pool.alloc(Bn, h, wd, cout)
            if cin == cout:
                d = K.conv3x3_halo_desc([h1], w[p + ".conv2.w"], out, gn=(coef2, True), bias=w[p + ".conv2.b"], residual=x)
            else:
                d = K.conv3x3_halo_desc([h1], w[p + ".conv2.w"], out, gn=(coef2, True), bias=w[p + ".conv2.b"],
                                        extra_srcs=[x], extra_segs=[(0, 0, 0, 0, K.ceil64(cin) // 64)])
            self._produce(d, out)
            pool.free(coef2); pool.free(h1); pool.free(x)
            return out
        cs = self._colstats(x, None)
        y = pool.alloc(Bn, h, wd, cin)
        prog.add(K.group_norm_op(x, None, w[p + ".gn1.g"], w[p + ".gn1.b"], y, self.stats, 1e-6, True, colstats=cs))
        h1 = pool.alloc(Bn, h, wd, cout)
        self._produce(K.conv3x3_desc(y, w[p + ".conv1.w"], h1, bias=w[p + ".conv1.b"]), h1)
        pool.free(y)
        cs = self._colstats(h1, None)
        y2 = pool.alloc(Bn, h, wd, cout)
        prog.add(K.group_norm_op(h1, None, w[p + ".gn2.g"], w[p + ".gn2.b"], y2, self.stats, 1e-6, True, colstats=cs))
        pool.free(h1)
        out = pool.alloc(Bn, h, wd, cout)
        nch = K.ceil64(cout) // 64
        segs = [(0, 0, kx - 1, ky - 1, nch) for ky in range(3) for kx in range(3)]
        if cin == cout:
            self._produce(K.gemm_desc([y2], segs, w[p + ".conv2.w"], out, cout, bias=w[p + ".conv2.b"], residual=x), out)
        else:
            segs.append((1, 0, 0, 0, K.ceil64(cin) // 64))
            self._produce(K.gemm_desc([y2, x], segs, w[p + ".conv2.w"], out, cout, bias=w[p + ".conv2.b"]), out)
        pool.free(y2); pool.free(x)
        return out

    def _attn(self, p, x):
        """AttnBlock (LD.py:3631-3642): single head, d = C = 512 -> GEMM / row-softmax / GEMM per image.
        The v bias is added after P@V (rows of P sum to 1), which lets V^T come straight out of a GEMM."""
        w, pool, prog = self.eng.w, self.pool, self.prog
        Bn, h, wd, c = x.shape
        n = h * wd
        cs = self._colstats(x, None)
        y = pool.alloc(Bn, h, wd, c)
        prog.add(K.group_norm_op(x, None, w[p + ".gn.g"], w[p + ".gn.b"], y, self.stats, 1e-6, False, colstats=cs))
        qk = pool.alloc(Bn * n, 2 * c)
        prog.add(K.linear_desc(y, w[p + ".qk.w"], qk, bias=w[p + ".qk.b"]))
        # P V contracts over the n tokens: its operands' row pitch is padded to a multiple of 64 (the GEMM's K chunk),
        # with V^T's padding columns zero, so that token counts off the 64 grid (odd image sizes) work too
        n64 = K.ceil64(n)
        if n64 == n:
            vt = pool.alloc(Bn, c, n)
        else:
            vt = torch.zeros(Bn, c, n64, dtype=F16, device=x.device)
            prog.keep(vt)
        s = pool.alloc(n, n, dtype=F32)
        pm = pool.alloc(n, n64)
        o = pool.alloc(Bn * n, c)
        y2, qk3, o3 = y.view(Bn, n, c), qk.view(Bn, n, 2 * c), o.view(Bn, n, c)
        for b in range(Bn):
            prog.add(K.linear_desc(w[p + ".v.w"], y2[b], vt[b, :, :n]))                # V^T = Wv @ hn^T
            prog.add(K.linear_desc(qk3[b, :, :c], qk3[b, :, c:], s, flags=L.EPI_OUT_F32))  # S = Q K^T
            prog.add(K.generic_op(L.OP_SOFTMAX, p=[s.data_ptr(), pm.data_ptr()], l=[n, n64, n], i=[n],
                                  f=[1.0 / math.sqrt(c)]))       # l: lds, ldo, rows
            prog.add(K.linear_desc(pm[:, :n], vt[b], o3[b], bias=w[p + ".v.b"]))      # O = P V + b_v
        out = pool.alloc(Bn, h, wd, c)
        self._produce(K.linear_desc(o, w[p + ".o.w"], out, bias=w[p + ".o.b"], residual=x), out)
        for t in (y, qk, s, pm, o, x) + ((vt,) if n64 == n else ()):
            pool.free(t)
        return out

    def run(self, z) -> torch.Tensor:
        self.z_in.copy_(z)
        if not self.use_graph:
            self.prog.run()
            return self.img
        if self.graph is None:
            self.prog.run()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.prog.run()
            self.graph = g
        self.graph.replay()
        return self.img


# =============================================================================== VAE encoder
class VAEEncoderEngine(VAEDecoderEngine):
    """AutoencoderKL encoder (Encoder + quant_conv + DiagonalGaussianRegularizer, LD.py:3649-3758, 3461-3481,
    3446-3458) with reference parameter names; shares the ResnetBlock / AttnBlock plans of the decoder."""

    def __init__(self, state_dict, cfg: VAEConfig = SD15_VAE, device="cuda"):
        L.require_device()
        self.cfg, self.dev = cfg, torch.device(device)
        self.w: Dict[str, torch.Tensor] = {}
        self._plans = _PlanCache()
        sd, dev, w = state_dict, self.dev, self.w
        w["cin.w"] = K.pack_conv3x3(UNetEngine._pad_in(sd["encoder.conv_in.weight"]))
        w["cin.b"] = _f32(sd, "encoder.conv_in.bias", dev)
        for i_level, blocks, has_down in vae_encoder_levels(cfg):
            for j in range(len(blocks)):
                self._pack_res(sd, f"encoder.down.{i_level}.block.{j}")
            if has_down:
                n = f"encoder.down.{i_level}.downsample.conv"
                w[n + ".w"], w[n + ".b"] = K.pack_conv3x3(sd[n + ".weight"]), _f32(sd, n + ".bias", dev)
        self._pack_res(sd, "encoder.mid.block_1")
        self._pack_res(sd, "encoder.mid.block_2")
        a = "encoder.mid.attn_1"
        w[a + ".gn.g"], w[a + ".gn.b"] = _f32(sd, a + ".norm.weight", dev), _f32(sd, a + ".norm.bias", dev)
        w[a + ".qk.w"] = K.pack_linear(torch.cat([sd[a + ".q.weight"], sd[a + ".k.weight"]], 0))
        w[a + ".qk.b"] = torch.cat([_f32(sd, a + ".q.bias", dev), _f32(sd, a + ".k.bias", dev)]).contiguous()
        w[a + ".v.w"], w[a + ".v.b"] = K.pack_linear(sd[a + ".v.weight"]), _f32(sd, a + ".v.bias", dev)
        w[a + ".o.w"], w[a + ".o.b"] = K.pack_linear(sd[a + ".proj_out.weight"]), _f32(sd, a + ".proj_out.bias", dev)
        w["nout.g"], w["nout.b"] = _f32(sd, "encoder.norm_out.weight", dev), _f32(sd, "encoder.norm_out.bias", dev)
        w["cout.w"], w["cout.b"] = K.pack_conv3x3(sd["encoder.conv_out.weight"]), _f32(sd, "encoder.conv_out.bias", dev)
        w["q.w"], w["q.b"] = K.pack_linear(sd["quant_conv.weight"]), _f32(sd, "quant_conv.bias", dev)
        torch.cuda.synchronize()

    def plan(self, B, H, W) -> "VAEEncodePlan":
        return self._plans.get((B, H, W), lambda: VAEEncodePlan(self, B, H, W))

    def decode(self, z):
        raise L.LdxError("VAEEncoderEngine has no decoder; use VAEDecoderEngine")

    decode_raw = decode

    @torch.no_grad()
    def encode(self, pixels: torch.Tensor, noise: Optional[torch.Tensor] = None, sample: bool = True) -> torch.Tensor:
        """VAE.encode (LD.py:6383-6410): pixels fp32 [B, H, W, 3] in [0,1] -> latent fp32 [B, 4, H/8, W/8].
        The posterior noise defaults to torch.randn(mean.shape) from the global CPU generator, the draw
        DiagonalGaussianDistribution.sample makes (LD.py:175-179); ``sample=False`` returns the posterior mean."""
        if pixels.dim() != 4 or pixels.shape[-1] < 3:
            raise L.LdxError(f"pixels must be [B, H, W, 3], got {tuple(pixels.shape)}")
        B, H, W, _ = pixels.shape
        f = 2 ** (len(self.cfg.ch_mult) - 1)
        if H % f or W % f:
            raise L.LdxError(f"image size must be a multiple of {f}, got {H}x{W}")
        pl = self.plan(B, H, W)
        if sample and noise is None:
            noise = torch.randn(B, self.cfg.z_channels, H // f, W // f)
        return pl.run(pixels[..., :3], noise if sample else None).clone()

    @torch.no_grad()
    def moments(self, pixels: torch.Tensor) -> torch.Tensor:
        """quant_conv(encoder(x)) as fp32 NCHW [B, 8, H/8, W/8] (mean | logvar), for tests."""
        B, H, W, _ = pixels.shape
        pl = self.plan(B, H, W)
        pl.run(pixels[..., :3], None)
        return pl.mom.permute(0, 3, 1, 2).contiguous()


class VAEEncodePlan(VAEPlan):
    def __init__(self, eng: VAEEncoderEngine, B, H, W):
        self.eng, self.B, self.H, self.W = eng, B, H, W
        dev, cfg, w = eng.dev, eng.cfg, eng.w
        self.pool = pool = _Pool(dev)
        self.prog = prog = K.Program()
        self._init_gn_stats()
        f = 2 ** (len(cfg.ch_mult) - 1)
        zc = cfg.z_channels
        self.pix_in = torch.zeros(B, H, W, 3, dtype=F32, device=dev)
        self.noise_in = torch.zeros(B, zc, H // f, W // f, dtype=F32, device=dev)
        self.z = torch.zeros(B, zc, H // f, W // f, dtype=F32, device=dev)
        self.stats = torch.empty(B * L.GN_STATS_FLOATS_PER_BATCH, dtype=F32, device=dev)
        x0 = torch.zeros(B, H, W, 8, dtype=F16, device=dev)
        prog.add(K.generic_op(L.OP_VAE_PRE, p=[self.pix_in.data_ptr(), x0.data_ptr()], l=[B * H * W]))
        h = pool.alloc(B, H, W, cfg.ch)
        self._produce(K.conv3x3_desc(x0, w["cin.w"], h, bias=w["cin.b"]), h)
        prog.keep(x0)
        for i_level, blocks, has_down in vae_encoder_levels(cfg):
            for j, (cin, cout) in enumerate(blocks):
                h = self._res(f"encoder.down.{i_level}.block.{j}", h, cin, cout)
            if has_down:
                n = f"encoder.down.{i_level}.downsample.conv"
                Bn, hh, ww, c = h.shape
                nh = pool.alloc(Bn, hh // 2, ww // 2, c)
                self._produce(K.conv3x3_s2_desc(h, w[n + ".w"], nh, pad_end=True, bias=w[n + ".b"]), nh)
                pool.free(h)
                h = nh
        top = cfg.ch * cfg.ch_mult[-1]
        h = self._res("encoder.mid.block_1", h, top, top)
        h = self._attn("encoder.mid.attn_1", h)
        h = self._res("encoder.mid.block_2", h, top, top)
        Bn, hh, ww, c = h.shape
        cs = self._colstats(h, None)
        y = pool.alloc(Bn, hh, ww, c)
        prog.add(K.group_norm_op(h, None, w["nout.g"], w["nout.b"], y, self.stats, 1e-6, True, colstats=cs))
        m0 = torch.zeros(Bn, hh, ww, 2 * zc, dtype=F16, device=dev)
        prog.add(K.conv3x3_desc(y, w["cout.w"], m0, bias=w["cout.b"]))
        self.mom = torch.zeros(Bn, hh, ww, 2 * zc, dtype=F32, device=dev)       # mean | logvar, fp32
        prog.add(K.linear_desc(m0, w["q.w"], self.mom, bias=w["q.b"], flags=L.EPI_OUT_F32))
        prog.keep(m0)
        self.mean_prog = K.Program()
        self.mean_prog.add(K.generic_op(L.OP_VAE_SAMPLE, p=[self.mom.data_ptr(), None, self.z.data_ptr()],
                                        l=[self.z.numel(), hh * ww], i=[zc]))
        self.sample_prog = K.Program()
        self.sample_prog.add(K.generic_op(L.OP_VAE_SAMPLE, p=[self.mom.data_ptr(), self.noise_in.data_ptr(),
                                                              self.z.data_ptr()],
                                          l=[self.z.numel(), hh * ww], i=[zc]))
        self.graph = None
        self.use_graph = True

    def run(self, pixels, noise=None) -> torch.Tensor:
        """Returns the plan's latent buffer (fp32 NCHW); valid until the next run."""
        self.pix_in.copy_(pixels)
        if noise is not None:
            self.noise_in.copy_(noise)
        if not self.use_graph:
            self.prog.run()
        else:
            if self.graph is None:
                self.prog.run()
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.prog.run()
                self.graph = g
            self.graph.replay()
        (self.sample_prog if noise is not None else self.mean_prog).run()
        return self.z
