"""Device time of the 23-block 4x RRDBNet on one 512x512 tile (the tile size of ImageUpscaleWithModel, LD.py:7369)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lightdiffusion_b200 import sd15
from lightdiffusion_b200.esrgan import ESRGANEngine, esrgan_param_shapes
dev = torch.device("cuda", 0)
eng = ESRGANEngine(sd15.seeded_state_dict(esrgan_param_shapes(23, 4), 303, dtype=torch.float16, device=dev))
for (b, h, w) in ((1, 256, 256), (1, 512, 512)):
    x = torch.rand(b, 3, h, w, device=dev)
    pl = eng.plan(b, h, w)
    eng(x); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): pl.prog.run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    fl = sum(m[1] for m in pl.prog.meta)
    print(f"esrgan x4 {b}x{h}x{w}: {ms:.2f} ms, {pl.prog.num_launches} launches, {fl/ms/1e9:.0f} TF/s executed (K padded to 64 per tap)", flush=True)
