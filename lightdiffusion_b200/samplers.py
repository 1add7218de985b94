"""k-diffusion sampler loops with the reference's signatures (seam #3):

    sample_X(model, x, sigmas, extra_args=None, callback=None, disable=None, eta=1.0, s_noise=1.0,
             noise_sampler=None)

``model(x, sigma_vec, **extra_args) -> denoised`` is the CFG guider.  The per-step state update runs
in one fused CUDA kernel (ldx_dpmpp2m_step / ldx_euler_ancestral_step); the scalar coefficients are
computed on the host from the sigmas with the same fp32 tensor arithmetic the reference uses
(the sigmas are 0-d tensors there, LD.py:1209-1236, 929-936).
"""
from __future__ import annotations

import math

import torch

from . import _lib as L

__all__ = ["sample_dpmpp_2m", "sample_dpmpp_2m_sde", "sample_euler_ancestral", "sample_dpm_adaptive", "PIDStepSizeController",
           "get_ancestral_step", "get_sigmas_karras", "dpmpp2m_coefficients"]


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


def dpmpp2m_coefficients(sig: torch.Tensor, i: int, h_last, eta: float = 0.0):
    """(c_x, c_d, c_o, h) of step i: x <- c_x x + c_d D + c_o (D - D_old).  LD.py:1209-1226 (solver 'midpoint')."""
    t, s = -sig[i].log(), -sig[i + 1].log()
    h = s - t
    eta_h = eta * h
    c_x = sig[i + 1] / sig[i] * (-eta_h).exp()
    c_d = (-h - eta_h).expm1().neg()
    c_o = None
    if h_last is not None:
        r = h_last / h
        c_o = 0.5 * (-h - eta_h).expm1().neg() * (1 / r)
    return float(c_x), float(c_d), (None if c_o is None else float(c_o)), h


@torch.no_grad()
def sample_dpmpp_2m(model, x, sigmas, extra_args=None, callback=None, disable=None, eta=0.0, s_noise=1.0,
                    noise_sampler=None):
    """DPM-Solver++(2M): the reference's sample_dpmpp_2m_sde (LD.py:1174-1244), solver 'midpoint'.  The north-star
    sampler is its deterministic case eta = 0 (SURVEY gotcha 1), the default here; the reference's ksampler closure
    passes its own eta (LD.py:2800-2816).  With eta > 0 the SDE noise term (LD.py:1229-1236) is added from the
    caller's ``noise_sampler(sigma, sigma_next)``; the reference's default BrownianTreeNoiseSampler (torchsde) is
    not part of this path, so eta > 0 without a noise_sampler raises instead of silently dropping the noise."""
    extra_args = {} if extra_args is None else extra_args
    _check_state(x)
    if eta and noise_sampler is None:
        raise L.LdxError("sample_dpmpp_2m: eta > 0 needs a noise_sampler (the reference's Brownian-tree default is "
                         "outside the hot path); pass eta=0 for the deterministic DPM++ 2M")
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
            c_x, c_d, c_o, h = dpmpp2m_coefficients(sig, i, h_last if old_denoised is not None else None, eta)
            L.check(lib.ldx_dpmpp2m_step(x.data_ptr(), denoised.data_ptr(),
                                         None if c_o is None else old_denoised.data_ptr(), c_x, c_d,
                                         0.0 if c_o is None else c_o, n, _stream()), "ldx_dpmpp2m_step")
            if eta:
                nz = noise_sampler(sig[i], sig[i + 1]).to(x.device, torch.float32).contiguous()
                coef = float(sig[i + 1] * (-2 * eta * h).expm1().neg().sqrt() * s_noise)
                # x <- x + coef * noise  (ldx_dpm_combine computes x - c1 * eps)
                L.check(lib.ldx_dpm_combine(x.data_ptr(), nz.data_ptr(), None, -coef, 0.0, x.data_ptr(), n, _stream()),
                        "ldx_dpm_combine")
        if callback is not None:
            callback({"x": x, "i": i, "sigma": sig[i], "sigma_hat": sig[i], "denoised": denoised})
        old_denoised = denoised.clone()      # the model may hand back a buffer it reuses next call
        h_last = h
    return x


def sample_dpmpp_2m_sde(model, x, sigmas, extra_args=None, callback=None, disable=None, eta=1.0, s_noise=1.0,
                        noise_sampler=None, solver_type="midpoint"):
    """The name and full signature the reference's ksampler closure resolves at call time (LD.py:1174-1185,
    2800-2816): assign this over the module global to route ``ksampler("dpmpp_2m_sde", {"eta": 0.0})`` -- the
    reference's spelling of DPM++ 2M -- through the fused update kernel."""
    if solver_type != "midpoint":
        raise L.LdxError(f"sample_dpmpp_2m_sde: solver_type {solver_type!r} is not on the hot path (midpoint only)")
    return sample_dpmpp_2m(model, x, sigmas, extra_args=extra_args, callback=callback, disable=disable, eta=eta,
                           s_noise=s_noise, noise_sampler=noise_sampler)


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


class PIDStepSizeController:
    """Step-size controller of dpm_solver_adaptive (interface of LD.py:944-973: the constructor arguments, ``h`` and
    ``propose_step(error) -> accepted``).  Pure host scalar arithmetic; it has to reproduce the reference's accept / reject
    sequence and step sizes bit for bit, so the floating-point expression of the factor -- newest ** k0 * previous ** k1 *
    oldest ** k2, evaluated left to right, then 1 + atan(f - 1) -- is the reference's; the bookkeeping is a two-entry history
    of ACCEPTED inverse errors instead of the reference's three-slot list (tests/test_host_logic.py checks both against each
    other on random error sequences)."""

    __slots__ = ("h", "_k", "_threshold", "_eps", "_accepted")

    def __init__(self, h, pcoeff, icoeff, dcoeff, order=1, accept_safety=0.81, eps=1e-8):
        self.h = h
        # exponents of the newest, the last accepted and the one-before-last accepted inverse error
        self._k = ((pcoeff + icoeff + dcoeff) / order, -(pcoeff + 2 * dcoeff) / order, dcoeff / order)
        self._threshold = accept_safety
        self._eps = eps
        self._accepted = None              # (last, one before last); seeded with the first proposal's own inverse error

    def propose_step(self, error) -> bool:
        newest = 1 / (float(error) + self._eps)
        if self._accepted is None:
            self._accepted = (newest, newest)
        last, before = self._accepted
        k0, k1, k2 = self._k
        factor = 1 + math.atan(newest ** k0 * last ** k1 * before ** k2 - 1)
        accepted = factor >= self._threshold
        if accepted:
            self._accepted = (newest, last)
        self.h *= factor
        return accepted


@torch.no_grad()
def sample_dpm_adaptive(model, x, sigma_min, sigma_max, extra_args=None, callback=None, disable=None, order=3,
                        rtol=0.05, atol=0.0078, h_init=0.05, pcoeff=0.0, icoeff=1.0, dcoeff=0.0, accept_safety=0.81,
                        eta=0.0, s_noise=1.0, noise_sampler=None, return_info=False):
    """DPM-Solver-12 and 23 with adaptive step size: sample_dpm_adaptive / DPMSolver.dpm_solver_adaptive
    (LD.py:1049-1170).  Per proposed step: three model calls (eps at s, s + h/3, s + 2h/3, shared by the embedded
    2nd- and 3rd-order updates through the reference's eps cache), four fused state updates, one fused error-norm
    reduction; the PID controller and every scalar coefficient run on the host with the reference's own 0-d fp32
    tensor arithmetic.  The reference never adds noise on this path (``su`` is fixed at 0.0, LD.py:1086), so
    ``eta`` only changes the controller's order, as it does there."""
    if sigma_min <= 0 or sigma_max <= 0:
        raise ValueError("sigma_min and sigma_max must not be 0")
    extra_args = {} if extra_args is None else extra_args
    _check_state(x)
    lib, n, st = L.lib(), x.numel(), _stream()
    sig_of = lambda t: t.neg().exp()
    t_start = -torch.tensor(float(sigma_max)).log()
    t_end = -torch.tensor(float(sigma_min)).log()
    forward = bool(t_end > t_start)
    h_init = abs(h_init) * (1 if forward else -1)
    s = t_start
    x = x.clone()
    x_prev = x.clone()
    ones = x.new_ones([x.shape[0]])
    eps, eps_r1, eps_r2, u, x_low, x_high = (torch.empty_like(x) for _ in range(6))
    ws = torch.empty(L.DPM_ERROR_WS_DOUBLES, dtype=torch.float64, device=x.device)
    err = torch.empty(1, dtype=torch.float32, device=x.device)
    pid = PIDStepSizeController(h_init, pcoeff, icoeff, dcoeff, 1.5 if eta else order, accept_safety)
    info = {"steps": 0, "nfe": 0, "n_accept": 0, "n_reject": 0}

    def model_eps(xx, t, out):
        sg = float(sig_of(t))
        den = model(xx, sg * ones, **extra_args)
        _check_state(den)
        L.check(lib.ldx_dpm_eps(xx.data_ptr(), den.data_ptr(), sg, out.data_ptr(), n, st), "ldx_dpm_eps")

    def combine(c1, e2, c2, out):
        L.check(lib.ldx_dpm_combine(x.data_ptr(), eps.data_ptr(), None if e2 is None else e2.data_ptr(), float(c1),
                                    0.0 if c2 is None else float(c2), out.data_ptr(), n, st), "ldx_dpm_combine")

    r1, r2 = 1 / 3, 2 / 3
    while (s < t_end - 1e-5) if forward else (s > t_end + 1e-5):
        t = torch.minimum(t_end, s + pid.h) if forward else torch.maximum(t_end, s + pid.h)
        h = t - s
        model_eps(x, s, eps)                                                       # eps(x, s)
        s1, s2 = s + r1 * h, s + r2 * h
        combine(sig_of(s1) * (r1 * h).expm1(), None, None, u)                      # u1 (shared by both orders)
        model_eps(u, s1, eps_r1)
        combine(sig_of(t) * h.expm1(), eps_r1, sig_of(t) / (2 * r1) * h.expm1(), x_low)       # x_2, LD.py:1009-1013
        combine(sig_of(s2) * (r2 * h).expm1(), eps_r1,
                sig_of(s2) * (r2 / r1) * ((r2 * h).expm1() / (r2 * h) - 1), u)                # u2, LD.py:1024-1031
        model_eps(u, s2, eps_r2)
        combine(sig_of(t) * h.expm1(), eps_r2, sig_of(t) / r2 * (h.expm1() / h - 1), x_high)  # x_3, LD.py:1033-1037
        L.check(lib.ldx_dpm_error_norm(x_low.data_ptr(), x_high.data_ptr(), x_prev.data_ptr(), float(torch.tensor(atol)),
                                       float(torch.tensor(rtol)), n, ws.data_ptr(), err.data_ptr(), st),
                "ldx_dpm_error_norm")
        accept = pid.propose_step(float(err))            # the one host sync of the step (float(error) in the reference)
        if accept:
            x_prev, x_low = x_low, x_prev
            x, x_high = x_high, x
            s = t
            info["n_accept"] += 1
        else:
            info["n_reject"] += 1
        info["nfe"] += order
        info["steps"] += 1
    # ``callback`` is accepted for signature compatibility: the reference stores it as info_callback (LD.py:1140-1147)
    # and dpm_solver_adaptive never invokes it
    if return_info:
        return x, info
    return x
