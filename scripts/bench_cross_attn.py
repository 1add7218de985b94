import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lightdiffusion_b200 import kernels as K
def timeit(fn, iters=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for (B, Nq, Nk, H, d) in [(16, 4096, 77, 8, 40), (16, 4096, 64, 8, 40), (16, 4096, 128, 8, 40), (16, 1024, 77, 8, 80), (16, 1024, 64, 8, 80)]:
    q = torch.randn(B, Nq, H * d, device="cuda").half(); k = torch.randn(B, Nk, H * d, device="cuda").half(); v = torch.randn(B, Nk, H * d, device="cuda").half()
    o = torch.empty_like(q)
    prog = K.Program(); prog.add(K.attention_op(q, k, v, o, H))
    ms = timeit(prog.run)
    print(f"attn B={B} Nq={Nq} Nk={Nk} d={d}: {ms*1e3:8.1f} us", flush=True)
