"""CLIP text tower of SD1.x on the CUDA path: CLIPTextModel (LD.py:4268-4486) as one recorded program per batch
size, and the prompt-weighting host logic around it (SDClipModel / ClipTokenWeightEncoder, LD.py:4539-4724).

Per layer: LayerNorm kernel -> fused q|k|v GEMM -> causal tcgen05 attention (d = 64) -> out_proj GEMM with the
residual fused -> LayerNorm -> fc1 GEMM with the quick_gelu epilogue -> fc2 GEMM with the residual fused.
The tokenizer (BPE vocabulary, prompt parser) stays with the caller: this module takes the token ids and weights
``SDTokenizer.tokenize_with_weights`` produces.  No CPU / torch fallback.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from . import kernels as K
from .sd15 import SD15_CLIP, CLIPTextConfig

F16, F32 = torch.float16, torch.float32


def gen_empty_tokens(special_tokens: dict, length: int) -> List[int]:
    """gen_empty_tokens, LD.py:4526-4536."""
    out = []
    if special_tokens.get("start") is not None:
        out.append(special_tokens["start"])
    if special_tokens.get("end") is not None:
        out.append(special_tokens["end"])
    return out + [special_tokens["pad"]] * (length - len(out))


class CLIPTextEngine:
    """CLIPTextModel with reference parameter names (``text_model.*``, ``text_projection.weight``)."""

    def __init__(self, state_dict, cfg: CLIPTextConfig = SD15_CLIP, device="cuda"):
        L.require_device()
        self.cfg, self.dev = cfg, torch.device(device)
        if cfg.hidden_size // cfg.num_heads != 64:
            raise L.LdxError("the causal attention kernel is built for 64-wide heads (CLIP ViT-L/14 text tower)")
        sd, dev = state_dict, self.dev
        f32 = lambda n: sd[n].to(device=dev, dtype=F16).float().contiguous()
        self.w: Dict[str, torch.Tensor] = {}
        w = self.w
        # the reference keeps the embedding tables in fp32 (LD.py:4423)
        w["tok"] = sd["text_model.embeddings.token_embedding.weight"].to(device=dev, dtype=F32).contiguous()
        w["pos"] = sd["text_model.embeddings.position_embedding.weight"].to(device=dev, dtype=F32).contiguous()
        for i in range(cfg.num_layers):
            p = f"text_model.encoder.layers.{i}"
            for ln in ("layer_norm1", "layer_norm2"):
                w[f"{i}.{ln}.g"], w[f"{i}.{ln}.b"] = f32(f"{p}.{ln}.weight"), f32(f"{p}.{ln}.bias")
            w[f"{i}.qkv.w"] = K.pack_linear(torch.cat([sd[f"{p}.self_attn.{n}_proj.weight"] for n in "qkv"], 0))
            w[f"{i}.qkv.b"] = torch.cat([f32(f"{p}.self_attn.{n}_proj.bias") for n in "qkv"]).contiguous()
            w[f"{i}.o.w"], w[f"{i}.o.b"] = K.pack_linear(sd[f"{p}.self_attn.out_proj.weight"]), f32(f"{p}.self_attn.out_proj.bias")
            w[f"{i}.fc1.w"], w[f"{i}.fc1.b"] = K.pack_linear(sd[f"{p}.mlp.fc1.weight"]), f32(f"{p}.mlp.fc1.bias")
            w[f"{i}.fc2.w"], w[f"{i}.fc2.b"] = K.pack_linear(sd[f"{p}.mlp.fc2.weight"]), f32(f"{p}.mlp.fc2.bias")
        w["fln.g"], w["fln.b"] = f32("text_model.final_layer_norm.weight"), f32("text_model.final_layer_norm.bias")
        w["proj.w"] = K.pack_linear(sd["text_projection.weight"])
        self._plans: Dict[Tuple[int, int, Optional[int]], "_CLIPPlan"] = {}
        torch.cuda.synchronize()

    def plan(self, B: int, T: int, intermediate: Optional[int]) -> "_CLIPPlan":
        key = (B, T, intermediate)
        if key not in self._plans:
            self._plans[key] = _CLIPPlan(self, B, T, intermediate)
        return self._plans[key]

    @torch.no_grad()
    def forward(self, input_tokens, intermediate_output: Optional[int] = None,
                final_layer_norm_intermediate: bool = True):
        """CLIPTextModel.forward (LD.py:4437-4486): returns (last, intermediate or None, projected pooled, pooled),
        fp32.  ``input_tokens`` is an integer [B, 77] tensor or nested list."""
        tok = torch.as_tensor(input_tokens)
        if tok.dim() != 2 or tok.shape[1] != self.cfg.max_positions:
            # the reference adds the whole position table (LD.py:4410), so any other length fails there too
            raise L.LdxError(f"tokens must be [B, {self.cfg.max_positions}], got {tuple(tok.shape)}")
        if int(tok.min()) < 0 or int(tok.max()) >= self.cfg.vocab_size:
            raise L.LdxError("token id outside the vocabulary (textual-inversion embeddings are not supported)")
        B, T = tok.shape
        inter = intermediate_output
        if inter is not None:
            inter = inter + self.cfg.num_layers if inter < 0 else inter
            if not 0 <= inter < self.cfg.num_layers:
                inter = None                      # the reference's loop never matches such an index
        pl = self.plan(B, T, inter)
        last, mid = pl.run(tok.to(self.dev, torch.int32), final_layer_norm_intermediate)
        eos = tok.to(self.dev, torch.int).argmax(dim=-1)       # pooled = the row of the largest id (EOS), LD.py:4451-4454
        pooled16 = last[torch.arange(B, device=self.dev), eos].contiguous()
        proj = torch.empty(B, self.cfg.hidden_size, dtype=F16, device=self.dev)
        K.run_gemm(K.linear_desc(pooled16, self.w["proj.w"], proj))
        return last.float(), None if mid is None else mid.float(), proj.float(), pooled16.float()


class _CLIPPlan:
    def __init__(self, eng: CLIPTextEngine, B: int, T: int, intermediate: Optional[int]):
        cfg, w, dev = eng.cfg, eng.w, eng.dev
        C, H, Fd = cfg.hidden_size, cfg.num_heads, cfg.intermediate_size
        self.prog = prog = K.Program()
        rows = B * T
        self.tokens = torch.zeros(B, T, dtype=torch.int32, device=dev)
        new = lambda *shape: torch.empty(*shape, dtype=F16, device=dev)
        x = new(rows, C)
        prog.add(K.generic_op(L.OP_CLIP_EMBED, p=[self.tokens.data_ptr(), w["tok"].data_ptr(), w["pos"].data_ptr(),
                                                  x.data_ptr()], l=[C], i=[rows, T, C]))
        y, qkv, att, hid = new(rows, C), new(B, T, 3 * C), new(B, T, C), new(rows, Fd)
        # every residual-stream tensor gets its own buffer (0.35 MB each at B = 3): nothing to alias, and the saved
        # intermediate (clip skip) stays valid
        keep = [x, y, qkv, att, hid]
        xin = x
        self.inter_raw = None
        for i in range(cfg.num_layers):
            prog.add(K.layer_norm_op(xin, w[f"{i}.layer_norm1.g"], w[f"{i}.layer_norm1.b"], y))
            prog.add(K.linear_desc(y, w[f"{i}.qkv.w"], qkv, bias=w[f"{i}.qkv.b"]))
            prog.add(K.attention_op(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], att, H, causal=True))
            xmid, xout = new(rows, C), new(rows, C)
            prog.add(K.linear_desc(att, w[f"{i}.o.w"], xmid, bias=w[f"{i}.o.b"], residual=xin))
            prog.add(K.layer_norm_op(xmid, w[f"{i}.layer_norm2.g"], w[f"{i}.layer_norm2.b"], y))
            prog.add(K.linear_desc(y, w[f"{i}.fc1.w"], hid, bias=w[f"{i}.fc1.b"], flags=L.EPI_QUICK_GELU))
            prog.add(K.linear_desc(hid, w[f"{i}.fc2.w"], xout, bias=w[f"{i}.fc2.b"], residual=xmid))
            keep += [xmid, xout]
            if i == intermediate:
                self.inter_raw = xout
            xin = xout
        self.last_raw = xin
        self.last = new(B, T, C)
        prog.add(K.layer_norm_op(self.last_raw, w["fln.g"], w["fln.b"], self.last.view(rows, C)))
        self.inter_ln = None
        if self.inter_raw is not None:
            self.inter_ln = new(B, T, C)
            prog.add(K.layer_norm_op(self.inter_raw, w["fln.g"], w["fln.b"], self.inter_ln.view(rows, C)))
        prog.keep(*keep)
        self.B, self.T, self.C = B, T, C

    def run(self, tokens: torch.Tensor, final_layer_norm_intermediate: bool = True):
        self.tokens.copy_(tokens)
        self.prog.run()
        if self.inter_raw is None:
            return self.last, None
        return self.last, (self.inter_ln if final_layer_norm_intermediate else self.inter_raw.view(self.B, self.T, self.C))


class SD1ClipEncoder:
    """SDClipModel + ClipTokenWeightEncoder (LD.py:4539-4724) over a CLIPTextEngine: ``encode_token_weights`` takes the
    tokenizer's list of 77-entry (token, weight) sections and returns (cond [1, 77 * sections, 768], pooled [1, 768])."""

    special_tokens = {"start": 49406, "end": 49407, "pad": 49407}

    def __init__(self, engine: CLIPTextEngine, layer: str = "last", layer_idx: Optional[int] = None,
                 layer_norm_hidden_state: bool = True, return_projected_pooled: bool = True):
        assert layer in ("last", "pooled", "hidden")
        self.engine = engine
        self.layer, self.layer_idx = layer, layer_idx
        self.layer_norm_hidden_state = layer_norm_hidden_state
        self.return_projected_pooled = return_projected_pooled
        self.options_default = (layer, layer_idx, return_projected_pooled)

    def set_clip_options(self, options: dict):
        """LD.py:4627-4633; CLIPSetLastLayer(-2) arrives here as {"layer": -2}."""
        self.layer_idx = options.get("layer", self.layer_idx)
        self.return_projected_pooled = options.get("projected_pooled", self.return_projected_pooled)
        self.layer = "hidden"

    def reset_clip_options(self):
        self.layer, self.layer_idx, self.return_projected_pooled = self.options_default

    def encode(self, tokens: Sequence[Sequence[int]]):
        """SDClipModel.forward, LD.py:4689-4724."""
        for sec in tokens:
            if not all(isinstance(t, int) for t in sec):
                raise L.LdxError("textual-inversion embeddings inside the token list are not supported")
        last, inter, proj, pooled = self.engine.forward(torch.tensor(tokens, dtype=torch.int64), self.layer_idx,
                                                        self.layer_norm_hidden_state)
        z = last if self.layer == "last" else inter
        if z is None:
            raise L.LdxError(f"layer index {self.layer_idx} selects no layer")
        return z, (proj if self.return_projected_pooled else pooled)

    def encode_token_weights(self, token_weight_pairs):
        """ClipTokenWeightEncoder.encode_token_weights, LD.py:4540-4569."""
        to_encode = [[int(t) if not torch.is_tensor(t) or t.dim() == 0 else t for t, _ in sec] for sec in token_weight_pairs]
        max_len = max((len(s) for s in to_encode), default=0)
        has_weights = any(wt != 1.0 for sec in token_weight_pairs for _, wt in sec)
        sections = len(to_encode)
        if has_weights or sections == 0:
            to_encode.append(gen_empty_tokens(self.special_tokens, max_len))
        out, pooled = self.encode(to_encode)
        first_pooled = pooled[0:1]
        output = []
        for k in range(sections):
            z = out[k:k + 1]
            if has_weights:
                wts = torch.tensor([wt for _, wt in token_weight_pairs[k]], device=z.device, dtype=z.dtype).view(1, -1, 1)
                z = torch.where(wts != 1.0, (z - out[-1]) * wts + out[-1], z)
            output.append(z)
        return torch.cat(output, dim=-2), first_pooled
