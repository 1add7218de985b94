"""Latent upscale between the two passes of hires-fix: ``bislerp`` / ``common_upscale`` with the reference's
signatures (LD.py:429-523), called by LatentUpscale.upscale (LD.py:6639-6654) with ``width // 8, height // 8``.

The reference gathers and slerps on whatever device the latent lives on (CPU fp32 in its GUI flow); here both
axis passes are one libldx.so kernel each (``ldx_bislerp_axis``).  The per-axis tables (ratio, left index, right
index) are ``generate_bilinear_data`` (LD.py:464-485) restated on the host: they are ``length_new`` numbers.
"""
from __future__ import annotations

from functools import lru_cache

import torch
import torch.nn.functional as F

from . import _lib as L


@lru_cache(maxsize=64)
def _axis_table_cpu(length_old: int, length_new: int):
    """generate_bilinear_data (LD.py:464-485): bilinear resampling of arange(length_old) gives the source coordinate;
    its fraction is the slerp ratio, its floor the left neighbour; the right neighbour comes from resampling
    arange + 1 with the last entry clamped."""
    c1 = torch.arange(length_old, dtype=torch.float32).reshape(1, 1, 1, -1)
    c1 = F.interpolate(c1, size=(1, length_new), mode="bilinear")
    ratios = c1 - c1.floor()
    c2 = torch.arange(length_old, dtype=torch.float32).reshape(1, 1, 1, -1) + 1
    c2[:, :, :, -1] -= 1
    c2 = F.interpolate(c2, size=(1, length_new), mode="bilinear")
    return (ratios.reshape(-1).contiguous(), c1.to(torch.int64).reshape(-1).to(torch.int32).contiguous(),
            c2.to(torch.int64).reshape(-1).to(torch.int32).contiguous())


def axis_table(length_old: int, length_new: int, device):
    return tuple(t.to(device) for t in _axis_table_cpu(int(length_old), int(length_new)))


@torch.no_grad()
def bislerp(samples: torch.Tensor, width: int, height: int) -> torch.Tensor:
    """samples [n, c, h, w] -> [n, c, height, width]: slerp along the width, then along the height (LD.py:487-518)."""
    L.require_device()
    if samples.dim() != 4:
        raise L.LdxError(f"bislerp expects [n, c, h, w], got {tuple(samples.shape)}")
    if not samples.is_cuda:
        raise L.LdxError("bislerp: samples must live on the CUDA device (no CPU fallback)")
    n, c, h, w = samples.shape
    width, height = int(width), int(height)
    if width < 1 or height < 1:
        raise L.LdxError(f"bislerp: bad target size {width}x{height}")
    x = samples.float().contiguous()
    st = torch.cuda.current_stream().cuda_stream
    r, i1, i2 = axis_table(w, width, x.device)
    mid = torch.empty(n, c, h, width, dtype=torch.float32, device=x.device)
    L.check(L.lib().ldx_bislerp_axis(x.data_ptr(), mid.data_ptr(), r.data_ptr(), i1.data_ptr(), i2.data_ptr(),
                                     n, c, h, w, width, 1, st), "ldx_bislerp_axis")
    r2, j1, j2 = axis_table(h, height, x.device)
    out = torch.empty(n, c, height, width, dtype=torch.float32, device=x.device)
    L.check(L.lib().ldx_bislerp_axis(mid.data_ptr(), out.data_ptr(), r2.data_ptr(), j1.data_ptr(), j2.data_ptr(),
                                     n, c, 1, h, height, width, st), "ldx_bislerp_axis")
    return out.to(samples.dtype)


def common_upscale(samples, width, height, upscale_method="bislerp", crop="disabled"):
    """common_upscale (LD.py:521-523): the reference ignores ``upscale_method`` and ``crop`` and always bislerps."""
    return bislerp(samples, width, height)
