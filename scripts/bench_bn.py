import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lightdiffusion_b200 import _lib as L, kernels as K
dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
def timeit(fn, iters=7):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]
for (M, N, Kd) in [(65536, 320, 2880), (65536, 320, 320), (16384, 640, 5760), (16384, 640, 640), (4096, 1280, 11520), (4096, 1280, 1280), (65536, 256, 2304)]:
    a = torch.randn(M, Kd, device=dev).half(); w = (torch.randn(N, Kd, device=dev) * Kd ** -0.5).half()
    out = torch.empty(M, N, dtype=torch.float16, device=dev)
    line = f"{M}x{N}x{Kd}: "
    for bn in (64, 96, 128, 160, 192, 224, 256):
        d = K.linear_desc(a, w, out, bn=bn)
        ms = timeit(lambda: K.run_gemm(d))
        line += f" bn{bn}:{2.0*M*N*Kd/ms/1e9:6.0f}"
    print(line, flush=True)
