"""Drop-in replacements for the reference's operator surface.

Seam #1  ``operations`` namespace: ``Linear / Conv2d / GroupNorm / LayerNorm / conv_nd`` with the
         constructor signatures and parameter names of ``disable_weight_init`` (LD.py:2361-2415), so
         ``UNetModel1(..., operations=ldx_ops)`` builds and ``load_state_dict`` works unchanged.
Seam #2  ``optimized_attention(q, k, v, heads, mask=None)`` (LD.py:3966-3988).
Seam #4  ``UNetFunctionWrapper`` for ``ModelPatcher.set_model_unet_function_wrapper`` (LD.py:3277,
         called from calc_cond_batch LD.py:2558-2567): the whole-UNet fast path.

Modules take and return NCHW-indexed tensors like the reference, but keep memory channels-last, so a
chain of these ops never transposes.  There is no torch fallback: unsupported shapes raise.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib as L
from . import kernels as K

F16 = torch.float16


def _require(cond, msg):
    if not cond:
        raise L.LdxError(msg)


def _nhwc(x: torch.Tensor) -> torch.Tensor:
    """NCHW-indexed tensor -> [B,H,W,C] view with unit channel stride (copies only if it is not
    channels-last already)."""
    _require(x.is_cuda and x.dtype == F16, "ldx ops run on fp16 CUDA tensors only (no CPU fallback)")
    y = x.permute(0, 2, 3, 1)
    return y if y.is_contiguous() else y.contiguous()


class _Cached:
    """Packs weights once per (storage, version); repacks after load_state_dict / LoRA patching."""

    def _cached(self, key, tensor, fn):
        tag = (tensor.data_ptr(), tensor._version, tensor.dtype)
        c = self.__dict__.setdefault("_ldx_cache", {})
        if key not in c or c[key][0] != tag:
            c[key] = (tag, fn(tensor.detach()))
        return c[key][1]

    def _bias32(self):
        if self.bias is None:
            return None
        return self._cached("b", self.bias, lambda b: b.to(F16).float().contiguous())


class ldx_ops:
    """``operations`` namespace (same attribute names as the reference's disable_weight_init)."""

    class Linear(nn.Linear, _Cached):
        comfy_cast_weights = False
        weight_function = None
        bias_function = None

        def reset_parameters(self):
            return None

        def forward(self, x):
            _require(x.is_cuda and x.dtype == F16, "ldx Linear: fp16 CUDA input required")
            k = x.shape[-1]
            _require(k % 8 == 0, f"ldx Linear: in_features {k} must be a multiple of 8")
            a = x.reshape(-1, k)
            a = a if a.stride(-1) == 1 and a.stride(0) % 8 == 0 else a.contiguous()
            w = self._cached("w", self.weight, K.pack_linear)
            out = torch.empty(a.shape[0], self.out_features, dtype=F16, device=x.device)
            K.run_gemm(K.linear_desc(a, w, out, bias=self._bias32()))
            return out.view(*x.shape[:-1], self.out_features)

    class Conv2d(nn.Conv2d, _Cached):
        comfy_cast_weights = False
        weight_function = None
        bias_function = None

        def reset_parameters(self):
            return None

        def forward(self, x):
            ks, st, pd = self.kernel_size, self.stride, self.padding
            xh = _nhwc(x)
            b, h, w, cin = xh.shape
            _require(self.groups == 1 and self.dilation == (1, 1), "ldx Conv2d: groups/dilation unsupported")
            if cin % 8:           # first conv of a network (4 latent channels): pad the pixel pitch to 16 bytes
                xp = torch.zeros(b, h, w, (cin + 7) // 8 * 8, dtype=F16, device=x.device)
                xp[..., :cin] = xh
                xh = xp
            cout = self.out_channels
            bias = self._bias32()
            if ks == (1, 1) and st == (1, 1) and pd == (0, 0):
                wp = self._cached("w", self.weight, K.pack_linear)
                out = torch.empty(b, h, w, cout, dtype=F16, device=x.device)
                K.run_gemm(K.linear_desc(xh, wp, out, bias=bias))
            elif ks == (3, 3) and pd == (1, 1) and st == (1, 1):
                wp = self._cached("w", self.weight, K.pack_conv3x3)
                out = torch.empty(b, h, w, cout, dtype=F16, device=x.device)
                K.run_gemm(K.conv3x3_desc(xh, wp, out, bias=bias))
            elif ks == (3, 3) and pd == (1, 1) and st == (2, 2):
                wp = self._cached("w", self.weight, K.pack_conv3x3)
                out = torch.empty(b, h // 2, w // 2, cout, dtype=F16, device=x.device)
                K.run_gemm(K.conv3x3_s2_desc(xh, wp, out, bias=bias))
            else:
                raise L.LdxError(f"ldx Conv2d: kernel {ks} stride {st} padding {pd} is not on the SD1.5 hot path")
            return out.permute(0, 3, 1, 2)

    class GroupNorm(nn.GroupNorm, _Cached):
        comfy_cast_weights = False
        weight_function = None
        bias_function = None

        def reset_parameters(self):
            return None

        def forward(self, x):
            _require(self.num_groups == 32 and self.affine, "ldx GroupNorm: 32 affine groups only")
            xh = _nhwc(x)
            b = xh.shape[0]
            g = self._cached("g", self.weight, lambda t: t.to(F16).float().contiguous())
            bt = self._cached("b", self.bias, lambda t: t.to(F16).float().contiguous())
            y = torch.empty_like(xh)
            stats = torch.empty(b * L.GN_STATS_FLOATS_PER_BATCH, dtype=torch.float32, device=x.device)
            K.run_op(K.group_norm_op(xh, None, g, bt, y, stats, self.eps, False))
            return y.permute(0, 3, 1, 2)

    class LayerNorm(nn.LayerNorm, _Cached):
        comfy_cast_weights = False
        weight_function = None
        bias_function = None

        def reset_parameters(self):
            return None

        def forward(self, x):
            _require(x.is_cuda and x.dtype == F16 and len(self.normalized_shape) == 1, "ldx LayerNorm: fp16 CUDA, 1-D")
            c = x.shape[-1]
            a = x.reshape(-1, c)
            a = a if a.stride(-1) == 1 and a.stride(0) % 8 == 0 else a.contiguous()
            g = self._cached("g", self.weight, lambda t: t.to(F16).float().contiguous())
            bt = self._cached("b", self.bias, lambda t: t.to(F16).float().contiguous())
            y = torch.empty(a.shape[0], c, dtype=F16, device=x.device)
            K.run_op(K.layer_norm_op(a, g, bt, y, self.eps))
            return y.view(x.shape)

    @classmethod
    def conv_nd(cls, dims, *args, **kwargs):
        _require(dims == 2, "ldx conv_nd: 2-D only")
        return cls.Conv2d(*args, **kwargs)


def optimized_attention(q, k, v, heads, mask=None):
    """q [B,Nq,heads*d], k/v [B,Nk,heads*d] fp16 -> [B,Nq,heads*d]; no mask on this path."""
    _require(mask is None, "ldx optimized_attention: masks are not used on the SD1.5 hot path")

    def prep(t):
        ok = t.is_cuda and t.dtype == F16 and t.stride(-1) == 1 and t.stride(1) % 8 == 0 and \
            t.stride(0) == t.stride(1) * t.shape[1]
        return t if ok else t.contiguous()
    q, k, v = prep(q), prep(k), prep(v)
    o = torch.empty(q.shape, dtype=F16, device=q.device)
    K.run_op(K.attention_op(q, k, v, o, heads))
    return o


optimized_attention_masked = optimized_attention


class UNetFunctionWrapper:
    """model_function_wrapper: replaces BaseModel.apply_model for the batched [uncond, cond] call.

    ``wrapper(apply_model_fn, {"input", "timestep", "c", "cond_or_uncond"}) -> denoised`` (LD.py:2558-2567).

    calc_cond_batch hands over a FRESH ``torch.cat`` of the conditioning on every call (cond_cat -> CONDCrossAttn.concat,
    LD.py:647-663, 2546) and the caching allocator recycles addresses, so a tensor's (pointer, version, shape) says
    nothing about its content: the cross-attention K/V of the 16 blocks are recomputed from the context on every call
    (one [B * tokens, 768] x [768, 20480] GEMM, ~1 % of a UNet step).  Contexts of 77 * k tokens (long prompts, padded
    by the reference to the lcm of the cond / uncond lengths) select a plan of that context length."""

    def __init__(self, engine):
        self.engine = engine

    def __call__(self, apply_model_fn, params):
        x, sigma, c = params["input"], params["timestep"], params["c"]
        ctx = c["c_crossattn"]
        b, _, h, w = x.shape
        _require(ctx.dim() == 3 and ctx.shape[0] == b, f"c_crossattn must be [{b}, tokens, dim], got {tuple(ctx.shape)}")
        plan = self.engine.plan(b, h, w, ctx_len=ctx.shape[1])
        return plan.run(x.float(), sigma.float(), ctx).clone()

    def to(self, device):
        return self
