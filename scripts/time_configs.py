"""End-to-end device times of BASELINE.json configurations 4 and 5 (synthetic inputs, random-init SD1.5 weights):
  config 4: hires-fix = 20-step DPM++ 2M at 64x64 -> bislerp 2x -> 10-step/0.45 Euler-ancestral pass at 128x128 -> decode
  config 5: img2img at 1024x1024, batch 4 = VAE encode -> 10-step/0.45 Euler-ancestral pass at 128x128 -> VAE decode
Prints one JSON line per configuration (CUDA-event time of the third of three runs)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lightdiffusion_b200 import sd15
from lightdiffusion_b200.engine import UNetEngine, VAEDecoderEngine, VAEEncoderEngine
from lightdiffusion_b200.pipeline import Txt2ImgPipeline, calculate_sigmas, sigmas_for_denoise

dev = torch.device("cuda", 0)
mk = lambda shapes, seed: sd15.seeded_state_dict(shapes, seed, dtype=torch.float16, device=dev)
pipe = Txt2ImgPipeline(UNetEngine(mk(sd15.unet_param_shapes(), 1234), device=dev),
                       VAEDecoderEngine(mk(sd15.vae_decoder_param_shapes(), 99), device=dev),
                       VAEEncoderEngine(mk(sd15.vae_encoder_param_shapes(), 100), device=dev))
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)


def timed(fn):
    fn(); fn(); torch.cuda.synchronize()          # two untimed runs: plans are built on the first, their CUDA graphs captured on the second
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1), out


b = 2
cond, uncond = rn(b, 77, 768), rn(b, 77, 768)
ms, _ = timed(lambda: pipe.hires_fix(rn(b, 4, 64, 64), rn(b, 4, 128, 128), cond, uncond, 7.0, calculate_sigmas("karras", 20)))
n2 = len(sigmas_for_denoise("normal", 10, 0.45)) - 1
print(json.dumps({"config": "4: hires-fix 512->1024, batch %d" % b, "ms": round(ms, 1), "images_per_s": round(b / ms * 1e3, 3),
                  "unet_evals": {"64x64": 20, "128x128": n2}, "gpu_launches": pipe.gpu_launches}), flush=True)
b = 4
cond, uncond = rn(b, 77, 768), rn(b, 77, 768)
pix = torch.rand(b, 1024, 1024, 3, device=dev, generator=g)
ms, _ = timed(lambda: pipe.img2img(pix, rn(b, 4, 128, 128), cond, uncond, 8.0, 10, 0.45, posterior_noise=rn(b, 4, 128, 128)))
print(json.dumps({"config": "5: img2img 1024x1024, batch %d (encode + %d Euler-a steps + decode)" % (b, n2), "ms": round(ms, 1),
                  "images_per_s": round(b / ms * 1e3, 3), "gpu_launches": pipe.gpu_launches}), flush=True)
