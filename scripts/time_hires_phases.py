"""Device time of each phase of BASELINE config 4 (hires-fix, batch 2): base sample, bislerp, second pass, decode."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lightdiffusion_b200 import sd15
from lightdiffusion_b200.engine import UNetEngine, VAEDecoderEngine
from lightdiffusion_b200.pipeline import Txt2ImgPipeline, calculate_sigmas, sigmas_for_denoise
from lightdiffusion_b200.upscale import bislerp

dev = torch.device("cuda", 0)
mk = lambda shapes, seed: sd15.seeded_state_dict(shapes, seed, dtype=torch.float16, device=dev)
pipe = Txt2ImgPipeline(UNetEngine(mk(sd15.unet_param_shapes(), 1234), device=dev),
                       VAEDecoderEngine(mk(sd15.vae_decoder_param_shapes(), 99), device=dev))
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)
b = 2
cond, uncond = rn(b, 77, 768), rn(b, 77, 768)
n1, n2 = rn(b, 4, 64, 64), rn(b, 4, 128, 128)
sig1, sig2 = calculate_sigmas("karras", 20), sigmas_for_denoise("normal", 10, 0.45)


def timed(name, fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
    print(f"{name:28s} device {e0.elapsed_time(e1):8.1f} ms   wall {(time.perf_counter() - t0) * 1e3:8.1f} ms", flush=True)
    return out


for rep in range(2):
    print("run", rep)
    z1 = timed("sample 64x64, 20 steps", lambda: pipe.sample(n1, cond, uncond, 7.0, sig1, "dpmpp_2m"))
    up = timed("bislerp 2x", lambda: bislerp(z1, 128, 128))
    z2 = timed("sample 128x128, 10 steps", lambda: pipe.sample(n2, cond, uncond, 8.0, sig2, "euler_ancestral", up))
    img = timed("decode 1024x1024", lambda: pipe.vae.decode(z2))
    timed("hires_fix whole", lambda: pipe.hires_fix(n1, n2, cond, uncond, 7.0, sig1))
