"""Graph-replay time of the CFG UNet step (B = 2 * batch samples, 64x64 latent): what bench.py reports as unet_step_ms."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lightdiffusion_b200 import sd15
from lightdiffusion_b200.engine import UNetEngine
dev = torch.device("cuda", 0)
sd = sd15.seeded_state_dict(sd15.unet_param_shapes(), 1234, dtype=torch.float16, device=dev)
eng = UNetEngine(sd, device=dev)
b = 8
plan = eng.plan(2 * b, 64, 64, cfg_pairs=True)
plan.set_context(torch.randn(2 * b, 77, 768, device=dev))
x = torch.randn(b, 4, 64, 64, device=dev) * 10
s = torch.full((b,), 5.0, device=dev)
for _ in range(3): plan.run(x, s, None, cfg_scale=7.0)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): plan.run(plan.x_in, plan.sigma_in, None, cfg_scale=7.0)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 20)
print(f"unet_step_ms {best:.3f}")
