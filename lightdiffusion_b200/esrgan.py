"""ESRGAN / RRDBNet upscaler (LD.py:6839-7235) on the CUDA path: the model behind ImageUpscaleWithModel
(LD.py:7356-7395), which the reference's Ultimate-SD-Upscale flow runs tile by tile through ``tiled_scale``.

Every layer is a 3x3 convolution on the implicit-GEMM kernel.  A ResidualDenseBlock keeps x, x1..x4 in channel
slices of ONE [B, H, W, nf + 4*gc] buffer, so ``torch.cat((x, x1, ...), 1)`` (LD.py:6992-6996) is a channel-prefix
view and never copied; LeakyReLU(0.2) runs in the epilogue; ``x5 * 0.2 + x`` folds the 0.2 into conv5's packed
weights and adds x as the fused residual; the RRDB's ``out * 0.2 + x`` (LD.py:6902) rides the last conv5 as one more
K segment (0.2 * I over the block input) plus the fused residual.  The 2x nearest upsample + conv pairs use the
four-phase folded form of the UNet / VAE upsamplers.  Parameter names are the reference module's own
(``model.0.weight``, ``model.1.sub.{i}.RDB{j}.conv{k}.0.weight``, ...).  No CPU / torch fallback.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict

import torch

from . import _lib as L
from . import kernels as K
from .engine import _Pool

F16, F32 = torch.float16, torch.float32


def esrgan_param_shapes(num_blocks: int = 23, scale: int = 4, nf: int = 64, gc: int = 32, in_nc: int = 3,
                        out_nc: int = 3, arch: str = "realesrgan") -> "OrderedDict[str, tuple]":
    """name -> shape of an RRDBNet checkpoint.  ``arch``: "realesrgan" (conv_first, body.N.rdbJ.convK, conv_body,
    conv_upN, conv_hr, conv_last), "esrgan" (conv_first, RRDB_trunk.N.RDBJ.convK, trunk_conv, upconvN, HRconv,
    conv_last) -- the two layouts the reference's loader accepts (state_map, LD.py:7043-7054) -- or "old", the module's
    own names (model.0, model.1.sub.N.RDBJ.convK.0, ...) that both are mapped to."""
    names = {"realesrgan": ("conv_first", "body.{i}.rdb{j}.conv{k}", "conv_body", "conv_up{u}", "conv_hr", "conv_last"),
             "esrgan": ("conv_first", "RRDB_trunk.{i}.RDB{j}.conv{k}", "trunk_conv", "upconv{u}", "HRconv", "conv_last"),
             "old": ("model.0", "model.1.sub.{i}.RDB{j}.conv{k}.0", "model.1.sub.{nb}", "model.{m}", "model.{hr}",
                     "model.{last}")}[arch]
    out: "OrderedDict[str, tuple]" = OrderedDict()

    def conv(name, cin, cout):
        out[name + ".weight"] = (cout, cin, 3, 3)
        out[name + ".bias"] = (cout,)
    nup = int(math.log2(scale))
    conv(names[0], in_nc, nf)
    for i in range(num_blocks):
        for j in (1, 2, 3):
            for k in range(1, 5):
                conv(names[1].format(i=i, j=j, k=k), nf + (k - 1) * gc, gc)
            conv(names[1].format(i=i, j=j, k=5), nf + 4 * gc, nf)
    conv(names[2].format(nb=num_blocks), nf, nf)
    for u in range(1, nup + 1):
        conv(names[3].format(u=u, m=3 * u), nf, nf)
    conv(names[4].format(hr=3 * nup + 2), nf, nf)          # HR conv (+ LeakyReLU)
    conv(names[5].format(last=3 * nup + 4), nf, out_nc)    # last conv
    return out


def to_old_arch(sd) -> "OrderedDict[str, torch.Tensor]":
    """RRDBNet.new_to_old_arch (LD.py:7142-7199): map the "esrgan" / "realesrgan" checkpoint layouts onto the module's
    own parameter names; a dict already in that layout passes through."""
    import re
    if "model.0.weight" in sd:
        return OrderedDict(sd)
    out = OrderedDict()
    blocks = []
    for k, v in sd.items():
        m = re.fullmatch(r"(?:RRDB_trunk|body)\.(\d+)\.(?:RDB|rdb)(\d)\.conv(\d+)\.(weight|bias)", k)
        if m:
            out[f"model.1.sub.{m.group(1)}.RDB{m.group(2)}.conv{m.group(3)}.0.{m.group(4)}"] = v
            blocks.append(int(m.group(1)))
    if not blocks or "conv_first.weight" not in sd:
        raise L.LdxError("not an RRDBNet state dict (expected conv_first.* and body.N.rdbJ.convK.* / RRDB_trunk.N.RDBJ.convK.*)")
    nb = max(blocks) + 1
    max_up = 0
    for k, v in sd.items():
        kind = k.rsplit(".", 1)[-1]
        base = k[:-len(kind) - 1]
        if base == "conv_first":
            out[f"model.0.{kind}"] = v
        elif base in ("trunk_conv", "conv_body"):
            out[f"model.1.sub.{nb}.{kind}"] = v
        else:
            m = re.fullmatch(r"(?:upconv|conv_up)(\d)", base)
            if m:
                out[f"model.{int(m.group(1)) * 3}.{kind}"] = v
                max_up = max(max_up, int(m.group(1)) * 3)
    for k, v in sd.items():
        kind = k.rsplit(".", 1)[-1]
        base = k[:-len(kind) - 1]
        if base in ("HRconv", "conv_hr"):
            out[f"model.{max_up + 2}.{kind}"] = v
        elif base == "conv_last":
            out[f"model.{max_up + 4}.{kind}"] = v
    return out


class ESRGANEngine:
    """RRDBNet.forward (LD.py:7224-7225): x fp32 NCHW [B, 3, h, w] -> [B, 3, h * scale, w * scale]."""

    def __init__(self, state_dict, device="cuda"):
        L.require_device()
        self.dev = torch.device(device)
        sd = to_old_arch(state_dict.get("params_ema", state_dict))      # load_state_dict, LD.py:7245-7251
        blocks = [int(k.split(".")[3]) for k in sd if ".RDB1.conv1.0.weight" in k]
        if not blocks or "model.0.weight" not in sd:
            raise L.LdxError("ESRGANEngine expects the old-arch RRDBNet layout (model.0.weight, model.1.sub.N.RDBj.convk.0.*)")
        self.num_blocks = max(blocks) + 1
        self.nf = sd["model.0.weight"].shape[0]
        self.gc = sd["model.1.sub.0.RDB1.conv1.0.weight"].shape[0]
        self.in_nc = sd["model.0.weight"].shape[1]
        tops = sorted(int(k.split(".")[1]) for k in sd if k.count(".") == 2 and k.endswith(".weight") and k != "model.0.weight")
        self.scale = 2 ** (len(tops) - 2)                 # upconvs + HR conv + last conv (get_scale, LD.py:7201-7209)
        self.out_nc = sd[f"model.{tops[-1]}.weight"].shape[0]
        if self.nf != 64 or self.gc != 32 or self.in_nc > 8 or self.out_nc > 8:
            raise L.LdxError("ESRGANEngine supports nf = 64, gc = 32 and at most 8 image channels")
        dev, nf, gc = self.dev, self.nf, self.gc
        f32 = lambda t, s=1.0: (t.to(device=dev, dtype=F16).float() * s).contiguous()
        w: Dict[str, torch.Tensor] = {}
        self.w = w
        w0 = sd["model.0.weight"]
        if w0.shape[1] % 8:
            w0 = torch.cat([w0, w0.new_zeros(w0.shape[0], 8 - w0.shape[1] % 8, 3, 3)], 1)
        w["first.w"], w["first.b"] = K.pack_conv3x3(w0), f32(sd["model.0.bias"])
        eye = torch.eye(nf, device=dev)
        for i in range(self.num_blocks):
            for j in (1, 2, 3):
                p = f"model.1.sub.{i}.RDB{j}"
                for k in range(1, 5):
                    w[f"{p}.c{k}.w"], w[f"{p}.c{k}.b"] = K.pack_conv3x3(sd[f"{p}.conv{k}.0.weight"]), f32(sd[f"{p}.conv{k}.0.bias"])
                # x5 * 0.2 + x: the scale lives in the packed weights; on the block's last RDB also the RRDB's
                # out * 0.2 + x_rrdb = 0.04 * x5 + 0.2 * x_rdb3 + x_rrdb, with 0.2 * I appended as a 1x1 K segment
                s5 = 0.2 if j < 3 else 0.2 * 0.2
                w5 = K.pack_conv3x3(sd[f"{p}.conv5.0.weight"].to(dev).float() * s5)
                if j == 3:
                    w5 = torch.cat([w5, K.pack_linear(0.2 * eye)], 1).contiguous()
                w[f"{p}.c5.w"], w[f"{p}.c5.b"] = w5, f32(sd[f"{p}.conv5.0.bias"], s5)
        t = f"model.1.sub.{self.num_blocks}"
        w["trunk.w"], w["trunk.b"] = K.pack_conv3x3(sd[t + ".weight"]), f32(sd[t + ".bias"])
        self.up_names = [f"model.{n}" for n in tops[:-2]]
        for n in self.up_names:
            w[n + ".w"], w[n + ".b"] = K.pack_upconv3x3(sd[n + ".weight"]), f32(sd[n + ".bias"])
        hr, last = f"model.{tops[-2]}", f"model.{tops[-1]}"
        w["hr.w"], w["hr.b"] = K.pack_conv3x3(sd[hr + ".weight"]), f32(sd[hr + ".bias"])
        w["last.w"], w["last.b"] = K.pack_conv3x3(sd[last + ".weight"]), f32(sd[last + ".bias"])
        self._plans: Dict[tuple, "_ESRGANPlan"] = {}
        torch.cuda.synchronize()

    def plan(self, B, H, W) -> "_ESRGANPlan":
        key = (B, H, W)
        if key not in self._plans:
            if len(self._plans) >= 4:                 # tiled_scale produces a few distinct edge-tile shapes
                self._plans.pop(next(iter(self._plans)))
            self._plans[key] = _ESRGANPlan(self, B, H, W)
        return self._plans[key]

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 4 or x.shape[1] != self.in_nc:
            raise L.LdxError(f"expected [B, {self.in_nc}, h, w], got {tuple(x.shape)}")
        if not x.is_cuda:
            raise L.LdxError("ESRGANEngine: input must live on the CUDA device (no CPU fallback)")
        B, _, H, W = x.shape
        return self.plan(B, H, W).run(x)

    __call__ = forward


class _ESRGANPlan:
    def __init__(self, eng: ESRGANEngine, B, H, W):
        self.eng = eng
        w, dev, nf, gc = eng.w, eng.dev, eng.nf, eng.gc
        self.prog = prog = K.Program()
        self.pool = pool = _Pool(dev)
        self.x_in = torch.zeros(B, eng.in_nc, H, W, dtype=F32, device=dev)
        x0 = torch.zeros(B, H, W, 8, dtype=F16, device=dev)
        prog.add(K.generic_op(L.OP_NCHW_TO_NHWC, p=[self.x_in.data_ptr(), None, x0.data_ptr()], i=[B, eng.in_nc, H, W, 8]))
        cw = nf + 4 * gc
        fea = pool.alloc(B, H, W, nf)
        prog.add(K.conv3x3_desc(x0, w["first.w"], fea, bias=w["first.b"]))
        # dense-block buffers: channels [x | x1 | x2 | x3 | x4]; three rotate (this RDB's, the next one's, the RRDB input)
        bufs = [pool.alloc(B, H, W, cw) for _ in range(3)]
        cur = bufs[0]
        # the first conv runs a second time straight into the first buffer's x slice (3 input channels: cheaper than a
        # copy kernel); `fea` itself stays untouched for the ShortcutBlock add at the end
        prog.add(K.conv3x3_desc(x0, w["first.w"], cur[..., :nf], bias=w["first.b"]))
        prog.keep(x0)
        for i in range(eng.num_blocks):
            rrdb_in = cur
            for j in (1, 2, 3):
                p = f"model.1.sub.{i}.RDB{j}"
                for c in range(1, 5):
                    cin = nf + (c - 1) * gc
                    prog.add(K.conv3x3_auto_desc(cur[..., :cin], w[f"{p}.c{c}.w"], cur[..., cin:cin + gc], bias=w[f"{p}.c{c}.b"],
                                                 flags=L.EPI_LRELU))
                nxt = next(b for b in bufs if b is not cur and b is not rrdb_in)
                if j < 3:
                    prog.add(K.conv3x3_auto_desc(cur, w[f"{p}.c5.w"], nxt[..., :nf], bias=w[f"{p}.c5.b"], residual=cur[..., :nf]))
                else:
                    prog.add(K.conv3x3_auto_desc(cur, w[f"{p}.c5.w"], nxt[..., :nf], bias=w[f"{p}.c5.b"], residual=rrdb_in[..., :nf],
                                                 extra_srcs=[cur[..., :nf]], extra_segs=[(0, 0, 0, 0, 1)]))
                cur = nxt
        trunk = pool.alloc(B, H, W, nf)
        prog.add(K.conv3x3_auto_desc(cur[..., :nf], w["trunk.w"], trunk, bias=w["trunk.b"], residual=fea))   # ShortcutBlock
        h, hh, ww = trunk, H, W
        for n in eng.up_names:
            up = pool.alloc(B, hh * 2, ww * 2, nf)
            prog.add(K.upconv3x3_descs(h, w[n + ".w"], up, bias=w[n + ".b"], flags=L.EPI_LRELU))
            h, hh, ww = up, hh * 2, ww * 2
        hr = pool.alloc(B, hh, ww, nf)
        prog.add(K.conv3x3_auto_desc(h, w["hr.w"], hr, bias=w["hr.b"], flags=L.EPI_LRELU))
        self.out_buf = torch.zeros(B, hh, ww, 8, dtype=F16, device=dev)
        prog.add(K.conv3x3_desc(hr, w["last.w"], self.out_buf[..., :eng.out_nc], bias=w["last.b"]))

    def run(self, x: torch.Tensor) -> torch.Tensor:
        self.x_in.copy_(x)
        self.prog.run()
        return self.out_buf[..., :self.eng.out_nc].permute(0, 3, 1, 2).float().contiguous()
