"""TAESD (the tiny autoencoder the reference uses for latent previews, LD.py:688-759) on the CUDA path: the
``taesd_decoder`` / ``taesd_encoder`` nn.Sequential stacks as recorded programs of 3x3 convolutions with ReLU
epilogues; a Block's ``relu(conv(x) + skip(x))`` (LD.py:695-703, skip = Identity at 64 channels) is the fused residual
followed by the ReLU epilogue, ``nn.Upsample(2) + conv`` the four-phase folded upsample conv, the stride-2 convs the
parity-view conv.  Parameter names are the Sequentials' own (``1.weight``, ``3.conv.0.weight``, ...), i.e. the keys
of taesd_decoder.safetensors / taesd_encoder.safetensors.  No CPU / torch fallback.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict

import torch

from . import _lib as L
from . import kernels as K
from .engine import _Pool

F16, F32 = torch.float16, torch.float32
NF = 64

# (kind, index) in nn.Sequential order: Decoder2 / Encoder2, LD.py:705-723
DECODER_LAYOUT = [("conv_relu", 1), ("block", 3), ("block", 4), ("block", 5), ("upconv", 7), ("block", 8), ("block", 9),
                  ("block", 10), ("upconv", 12), ("block", 13), ("block", 14), ("block", 15), ("upconv", 17),
                  ("block", 18), ("conv_out", 19)]
ENCODER_LAYOUT = [("conv_in", 0), ("block", 1), ("down", 2), ("block", 3), ("block", 4), ("block", 5), ("down", 6),
                  ("block", 7), ("block", 8), ("block", 9), ("down", 10), ("block", 11), ("block", 12), ("block", 13),
                  ("conv_out", 14)]


def taesd_param_shapes(which: str = "decoder", latent_channels: int = 4) -> "OrderedDict[str, tuple]":
    out: "OrderedDict[str, tuple]" = OrderedDict()
    layout = DECODER_LAYOUT if which == "decoder" else ENCODER_LAYOUT
    cin_first, cout_last = (latent_channels, 3) if which == "decoder" else (3, latent_channels)
    for kind, i in layout:
        if kind == "block":
            for j in (0, 2, 4):
                out[f"{i}.conv.{j}.weight"] = (NF, NF, 3, 3)
                out[f"{i}.conv.{j}.bias"] = (NF,)
        elif kind in ("upconv", "down"):
            out[f"{i}.weight"] = (NF, NF, 3, 3)                       # bias=False
        elif kind in ("conv_relu", "conv_in"):
            out[f"{i}.weight"] = (NF, cin_first, 3, 3)
            out[f"{i}.bias"] = (NF,)
        else:
            out[f"{i}.weight"] = (cout_last, NF, 3, 3)
            out[f"{i}.bias"] = (cout_last,)
    return out


class _TAESDStack:
    def __init__(self, state_dict, which: str, device="cuda"):
        L.require_device()
        self.which, self.dev = which, torch.device(device)
        self.layout = DECODER_LAYOUT if which == "decoder" else ENCODER_LAYOUT
        sd, dev = state_dict, self.dev
        missing = [k for k in taesd_param_shapes(which) if k not in sd]
        if missing:
            raise L.LdxError(f"TAESD {which}: missing parameters {missing[:3]}...")
        f32 = lambda n: sd[n].to(device=dev, dtype=F16).float().contiguous()
        self.w: Dict[str, torch.Tensor] = {}
        for kind, i in self.layout:
            if kind == "block":
                for j in (0, 2, 4):
                    self.w[f"{i}.{j}.w"], self.w[f"{i}.{j}.b"] = K.pack_conv3x3(sd[f"{i}.conv.{j}.weight"]), f32(f"{i}.conv.{j}.bias")
            elif kind == "upconv":
                self.w[f"{i}.w"] = K.pack_upconv3x3(sd[f"{i}.weight"])
            elif kind == "down":
                self.w[f"{i}.w"] = K.pack_conv3x3(sd[f"{i}.weight"])
            else:
                wt = sd[f"{i}.weight"]
                if wt.shape[1] % 8:
                    wt = torch.cat([wt, wt.new_zeros(wt.shape[0], 8 - wt.shape[1] % 8, 3, 3)], 1)
                self.w[f"{i}.w"], self.w[f"{i}.b"] = K.pack_conv3x3(wt), f32(f"{i}.bias")
        self.cin = sd[f"{self.layout[0][1]}.weight"].shape[1]
        self.cout = sd[f"{self.layout[-1][1]}.weight"].shape[0]
        self._plans: Dict[tuple, "_TAESDPlan"] = {}
        torch.cuda.synchronize()

    def _run(self, x: torch.Tensor, mode: int) -> torch.Tensor:
        if x.dim() != 4 or x.shape[1] != self.cin:
            raise L.LdxError(f"TAESD {self.which}: expected [B, {self.cin}, h, w], got {tuple(x.shape)}")
        if not x.is_cuda:
            raise L.LdxError("TAESD: input must live on the CUDA device (no CPU fallback)")
        B, _, H, W = x.shape
        if self.which == "encoder" and (H % 8 or W % 8):
            raise L.LdxError("TAESD encoder: image size must be a multiple of 8")
        key = (B, H, W, mode)
        if key not in self._plans:
            if len(self._plans) >= 4:
                self._plans.pop(next(iter(self._plans)))
            self._plans[key] = _TAESDPlan(self, B, H, W, mode)
        return self._plans[key].run(x)


class _TAESDPlan:
    def __init__(self, st: _TAESDStack, B, H, W, mode):
        w, dev = st.w, st.dev
        self.prog = prog = K.Program()
        pool = _Pool(dev)
        self.x_in = torch.zeros(B, st.cin, H, W, dtype=F32, device=dev)
        x0 = torch.zeros(B, H, W, 8, dtype=F16, device=dev)
        prog.add(K.generic_op(L.OP_NCHW_TO_NHWC, p=[self.x_in.data_ptr(), None, x0.data_ptr()], i=[B, st.cin, H, W, 8, mode]))
        prog.keep(x0)
        h, hh, ww = x0, H, W
        for kind, i in st.layout:
            if kind in ("conv_relu", "conv_in"):
                o = pool.alloc(B, hh, ww, NF)
                prog.add(K.conv3x3_desc(h, w[f"{i}.w"], o, bias=w[f"{i}.b"], flags=L.EPI_RELU if kind == "conv_relu" else 0))
                h = o
            elif kind == "block":
                a, b_ = pool.alloc(B, hh, ww, NF), pool.alloc(B, hh, ww, NF)
                prog.add(K.conv3x3_auto_desc(h, w[f"{i}.0.w"], a, bias=w[f"{i}.0.b"], flags=L.EPI_RELU))
                prog.add(K.conv3x3_auto_desc(a, w[f"{i}.2.w"], b_, bias=w[f"{i}.2.b"], flags=L.EPI_RELU))
                prog.add(K.conv3x3_auto_desc(b_, w[f"{i}.4.w"], a, bias=w[f"{i}.4.b"], residual=h, flags=L.EPI_RELU))
                pool.free(b_); pool.free(h) if h is not x0 else None
                h = a
            elif kind == "upconv":
                o = pool.alloc(B, hh * 2, ww * 2, NF)
                prog.add(K.upconv3x3_descs(h, w[f"{i}.w"], o))
                pool.free(h)
                h, hh, ww = o, hh * 2, ww * 2
            elif kind == "down":
                o = pool.alloc(B, hh // 2, ww // 2, NF)
                prog.add(K.conv3x3_s2_desc(h, w[f"{i}.w"], o))
                pool.free(h)
                h, hh, ww = o, hh // 2, ww // 2
            else:
                self.out_buf = torch.zeros(B, hh, ww, 8, dtype=F16, device=dev)
                prog.add(K.conv3x3_desc(h, w[f"{i}.w"], self.out_buf[..., :st.cout], bias=w[f"{i}.b"]))
        self.cout = st.cout

    def run(self, x):
        self.x_in.copy_(x)
        self.prog.run()
        return self.out_buf[..., :self.cout].permute(0, 3, 1, 2).float().contiguous()


class TAESDEngine:
    """TAESD.decode / TAESD.encode (LD.py:748-759) with vae_shift = 0 and vae_scale = 1 (their initial values;
    the reference never trains or loads them)."""

    def __init__(self, decoder_sd=None, encoder_sd=None, device="cuda"):
        self.decoder = _TAESDStack(decoder_sd, "decoder", device) if decoder_sd is not None else None
        self.encoder = _TAESDStack(encoder_sd, "encoder", device) if encoder_sd is not None else None

    @torch.no_grad()
    def decode(self, latent: torch.Tensor) -> torch.Tensor:
        """latent [B, 4, h, w] -> image [B, 3, 8h, 8w] in [-1, 1] (decoder output .sub(0.5).mul(2), LD.py:752-753)."""
        if self.decoder is None:
            raise L.LdxError("TAESDEngine was built without decoder weights")
        return self.decoder._run(latent.float(), 1).sub_(0.5).mul_(2)

    @torch.no_grad()
    def encode(self, image: torch.Tensor) -> torch.Tensor:
        """image [B, 3, H, W] in [-1, 1] -> latent [B, 4, H/8, W/8] (encoder(x * 0.5 + 0.5), LD.py:756-759)."""
        if self.encoder is None:
            raise L.LdxError("TAESDEngine was built without encoder weights")
        return self.encoder._run(image.float(), 2)
