"""Halo-tile conv (no transform) against the implicit-GEMM conv for explicit N tiles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lightdiffusion_b200 import _lib as L, kernels as K
dev = "cuda"

def timeit(fn, iters=7, reps=20):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps)
    ts.sort()
    return ts[len(ts) // 2]

for (B, H, W, Cin, Cout) in [(16, 32, 32, 640, 640), (16, 16, 16, 1280, 1280), (16, 64, 64, 320, 320)]:
    x = torch.randn(B, H, W, Cin, device=dev).half()
    w = (torch.randn(Cout, Cin, 3, 3, device=dev) * (9 * Cin) ** -0.5).half()
    wp = K.pack_conv3x3(w)
    out = torch.empty(B, H, W, Cout, dtype=torch.float16, device=dev)
    flops = 2.0 * B * H * W * Cout * 9 * Cin
    for bn in (0, 128, 160, 192, 224, 256, 320, 384, 512):
        if bn > Cout and bn != 0:
            continue
        row = f"B{B} {H}x{W} {Cin}->{Cout} bn={bn:3d}:"
        for halo in (False, True):
            try:
                d = K.conv3x3_halo_desc([x], wp, out, bn=bn) if halo else K.conv3x3_desc(x, wp, out, bn=bn)
                p = K.Program(); p.add(d)
                t = timeit(p.run)
                row += f"  {'halo' if halo else 'gemm'} {t*1e3:7.1f} us {flops/(t*1e-3)/1e12:6.0f} TF"
            except Exception as e:
                row += f"  {'halo' if halo else 'gemm'} n/a"
        print(row, flush=True)
