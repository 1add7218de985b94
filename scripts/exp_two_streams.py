"""Experiment: one B=16 UNet plan vs two B=8 plans replayed concurrently on two streams (tail/ramp bubbles of
one half's kernels filled by the other half's)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lightdiffusion_b200 import sd15
from lightdiffusion_b200.engine import UNetEngine, UNetPlan
dev = torch.device("cuda", 0)
sd = sd15.seeded_state_dict(sd15.unet_param_shapes(), 1234, dtype=torch.float16, device=dev)
eng = UNetEngine(sd, device=dev)
def mk(B):
    p = UNetPlan(eng, B, 64, 64, True)
    p.set_context(torch.randn(B, 77, 768, device=dev))
    x = torch.randn(B // 2, 4, 64, 64, device=dev) * 10
    s = torch.full((B // 2,), 5.0, device=dev)
    p.run(x, s, None, cfg_scale=7.0)
    torch.cuda.synchronize()
    return p
def t(fn, n=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
p16 = mk(16)
print(f"one plan B=16:            {t(lambda: p16.graph.replay()):.3f} ms")
pa, pb = mk(8), mk(8)
print(f"two plans B=8 serial:     {t(lambda: (pa.graph.replay(), pb.graph.replay())):.3f} ms")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1): pa.graph.replay()
    with torch.cuda.stream(s2): pb.graph.replay()
    cur.wait_stream(s1); cur.wait_stream(s2)
print(f"two plans B=8 concurrent: {t(both):.3f} ms")
pc, pd = mk(8), mk(8)   # four quarter... not applicable (B=4 each)
