import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lightdiffusion_b200 import _lib as L, kernels as K
dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
def timeit(fn, iters=7, do_flush=True):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        if do_flush: flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]
for (M, N) in [(65536, 320), (16384, 320), (16384, 640), (4096, 1280)]:
    for Kd in (64, 128, 320, 640, 1280, 2560, 5120):
        a = torch.randn(M, Kd, device=dev).half(); w = (torch.randn(N, Kd, device=dev) * Kd ** -0.5).half()
        out = torch.empty(M, N, dtype=torch.float16, device=dev)
        d = K.linear_desc(a, w, out, bn=int(os.environ.get("BN", "0")))
        c = timeit(lambda: K.run_gemm(d)); h = timeit(lambda: K.run_gemm(d), do_flush=False)
        print(f"M={M} N={N} K={Kd:5d}: cold {c*1e3:7.1f} us  warm {h*1e3:7.1f} us  {2.0*M*N*Kd/h/1e9:7.0f} TF/s(warm)", flush=True)
