"""Sweeps the N tile width (and 1-CTA vs CTA-pair via LDX_GEMM_CTAS) for a few UNet GEMM shapes: how far is the
cost model's choice from the best configuration?  Steady-state replay timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lightdiffusion_b200 import _lib as L, kernels as K
dev = "cuda"
def timeit(fn, reps=20, rounds=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(rounds):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps)
    return sorted(ts)[len(ts) // 2]
shapes = [(65536, 320, 320), (16384, 640, 640), (4096, 1280, 1280), (16384, 640, 2560), (4096, 1280, 5120), (65536, 320, 1280), (65536, 960, 320)]
for (M, N, Kd) in shapes:
    a = torch.randn(M, Kd, device=dev).half(); w = (torch.randn(N, Kd, device=dev) * Kd ** -0.5).half()
    b = torch.randn(N, device=dev); r = torch.randn(M, N, device=dev).half()
    out = torch.empty(M, N, dtype=torch.float16, device=dev)
    res = []
    for bn in (0, 96, 128, 160, 192, 224, 256):
        try:
            prog = K.Program(); prog.add(K.linear_desc(a, w, out, bias=b, residual=r, bn=bn))
            res.append((bn, timeit(prog.run) * 1e3))
        except Exception as e:
            res.append((bn, float("nan")))
    print(f"ctas={os.environ.get('LDX_GEMM_CTAS','auto')} M={M} N={N} K={Kd}: " + "  ".join(f"bn{bn}:{t:.1f}" for bn, t in res), flush=True)
