"""How much of a UNet step is launch skeleton?  Graph replay of the recorded program with every GEMM-family kernel
doing no tiles (LDX_GEMM_DEBUG=3: prologue, TMEM allocation, griddepcontrol, teardown only) against the normal replay
and the per-op sum of the other kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lightdiffusion_b200 import sd15
from lightdiffusion_b200.engine import UNetEngine
dev = torch.device("cuda", 0)
sd = sd15.seeded_state_dict(sd15.unet_param_shapes(), 1234, dtype=torch.float16, device=dev)
eng = UNetEngine(sd, device=dev)
plan = eng.plan(16, 64, 64, cfg_pairs=True)
plan.use_graph = False
plan.set_context(torch.randn(16, 77, 768, device=dev))
plan.run(torch.randn(8, 4, 64, 64, device=dev) * 10, torch.full((8,), 5.0, device=dev), None, cfg_scale=7.0)
torch.cuda.synchronize()
rows = plan.prog.run_timed(); rows = plan.prog.run_timed()
agg = {}
for kind, fl, tag, ms in rows:
    a = agg.setdefault(kind, [0, 0.0]); a[0] += 1; a[1] += ms
print({k: (v[0], round(v[1], 3)) for k, v in agg.items()})
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    plan.prog.run()
for _ in range(3): g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): g.replay()
e1.record(); torch.cuda.synchronize()
print(f"graph replay: {e0.elapsed_time(e1) / 20:.3f} ms, {plan.prog.num_launches} launches")
