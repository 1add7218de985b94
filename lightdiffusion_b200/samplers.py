"""k-diffusion sampler loops with the reference's signatures (seam #3):

    sample_X(model, x, sigmas, extra_args=None, callback=None, disable=None, eta=1.0, s_noise=1.0,
             noise_sampler=None)

``model(x, sigma_vec, **extra_args) -> denoised`` is the CFG guider.  The per-step state update runs
in one fused CUDA kernel (ldx_dpmpp2m_step / ldx_euler_ancestral_step); the scalar coefficients are
computed on the host from the sigmas with the same fp32 tensor arithmetic the reference uses
(the sigmas are 0-d tensors there, LD.py:1209-1236, 929-936).
"""
from __future__ import annotations

import torch

from . import _lib as L

__all__ = ["sample_dpmpp_2m", "sample_euler_ancestral", "get_ancestral_step", "get_sigmas_karras",
           "dpmpp2m_coefficients"]


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _host_sigmas(sigmas) -> torch.Tensor:
    return torch.as_tensor(sigmas).detach().to("cpu", torch.float32)


def _check_state(x):
    if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
        raise L.LdxError("sampler state must be a contiguous fp32 CUDA tensor (the reference keeps x in fp32)")


def get_sigmas_karras(n, sigma_min, sigma_max, rho=7.0, device="cpu"):
    """LD.py:831-837."""
    ramp = torch.linspace(0, 1, n, device=device)
    min_inv_rho, max_inv_rho = sigma_min ** (1 / rho), sigma_max ** (1 / rho)
    sigmas = (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho
    return torch.cat([sigmas, sigmas.new_zeros([1])]).to(device)


def get_ancestral_step(sigma_from, sigma_to, eta=1.0):
    """LD.py:844-850."""
    sigma_up = min(sigma_to, eta * (sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2) ** 0.5)
    sigma_down = (sigma_to ** 2 - sigma_up ** 2) ** 0.5
    return sigma_down, sigma_up


def dpmpp2m_coefficients(sig: torch.Tensor, i: int, h_last):
    """(c_x, c_d, c_o, h) of step i: x <- c_x x + c_d D + c_o (D - D_old).  LD.py:1209-1226 with eta = 0."""
    t, s = -sig[i].log(), -sig[i + 1].log()
    h = s - t
    eta_h = 0.0 * h
    c_x = sig[i + 1] / sig[i] * (-eta_h).exp()
    c_d = (-h - eta_h).expm1().neg()
    c_o = None
    if h_last is not None:
        r = h_last / h
        c_o = 0.5 * (-h - eta_h).expm1().neg() * (1 / r)
    return float(c_x), float(c_d), (None if c_o is None else float(c_o)), h


@torch.no_grad()
def sample_dpmpp_2m(model, x, sigmas, extra_args=None, callback=None, disable=None, eta=1.0, s_noise=1.0,
                    noise_sampler=None):
    """DPM-Solver++(2M): the reference's sample_dpmpp_2m_sde (LD.py:1174-1244) at eta=0, solver
    'midpoint' (SURVEY gotcha 1).  ``eta``/``s_noise``/``noise_sampler`` are accepted for signature
    compatibility; the deterministic sampler draws no noise."""
    extra_args = {} if extra_args is None else extra_args
    _check_state(x)
    sig = _host_sigmas(sigmas)
    x = x.clone()
    s_in = x.new_ones([x.shape[0]])
    old_denoised, h_last = None, None
    lib, n = L.lib(), x.numel()
    for i in range(len(sig) - 1):
        denoised = model(x, float(sig[i]) * s_in, **extra_args)
        _check_state(denoised)
        if sig[i + 1] == 0:
            x = denoised.clone()
            h = None
        else:
            c_x, c_d, c_o, h = dpmpp2m_coefficients(sig, i, h_last if old_denoised is not None else None)
            L.check(lib.ldx_dpmpp2m_step(x.data_ptr(), denoised.data_ptr(),
                                         None if c_o is None else old_denoised.data_ptr(), c_x, c_d,
                                         0.0 if c_o is None else c_o, n, _stream()), "ldx_dpmpp2m_step")
        if callback is not None:
            callback({"x": x, "i": i, "sigma": sig[i], "sigma_hat": sig[i], "denoised": denoised})
        old_denoised = denoised.clone()      # the model may hand back a buffer it reuses next call
        h_last = h
    return x


@torch.no_grad()
def sample_euler_ancestral(model, x, sigmas, extra_args=None, callback=None, disable=None, eta=1.0, s_noise=1.0,
                           noise_sampler=None):
    """LD.py:907-941.  The default noise sampler is torch.randn_like(x), as in the reference."""
    extra_args = {} if extra_args is None else extra_args
    _check_state(x)
    noise_sampler = (lambda s, sn: torch.randn_like(x)) if noise_sampler is None else noise_sampler
    sig = _host_sigmas(sigmas)
    x = x.clone()
    s_in = x.new_ones([x.shape[0]])
    lib, n = L.lib(), x.numel()
    for i in range(len(sig) - 1):
        denoised = model(x, float(sig[i]) * s_in, **extra_args)
        _check_state(denoised)
        sigma_down, sigma_up = get_ancestral_step(sig[i], sig[i + 1], eta=eta)
        noise = None
        if sig[i + 1] > 0:
            noise = noise_sampler(sig[i], sig[i + 1]).to(torch.float32).contiguous()
        L.check(lib.ldx_euler_ancestral_step(x.data_ptr(), denoised.data_ptr(),
                                             None if noise is None else noise.data_ptr(), float(sig[i]),
                                             float(sigma_down), float(s_noise * sigma_up), n, _stream()),
                "ldx_euler_ancestral_step")
        if callback is not None:
            callback({"x": x, "i": i, "sigma": sig[i], "sigma_hat": sig[i], "denoised": denoised})
    return x
