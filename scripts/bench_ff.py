"""Feed-forward of the 64 x 64-level transformer block (C = 320, M = 65536): GEGLU GEMM + second GEMM against the fused
kernel (ldx_ff_geglu_fused).  Steady state, 20 back-to-back replays."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lightdiffusion_b200 import _lib as L, kernels as K
dev = "cuda"
def timeit(fn, iters=7, reps=20):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / reps)
    ts.sort(); return ts[len(ts) // 2]
for M in (65536, 32768):
    C, Fd = 320, 1280
    x = (torch.randn(M, C, device=dev) * 1.2 + 0.3).half()
    gamma, beta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.2
    w1, b1 = (torch.randn(2 * Fd, C, device=dev) * C ** -0.5).half(), torch.randn(2 * Fd, device=dev) * 0.3
    w2, b2 = (torch.randn(C, Fd, device=dev) * Fd ** -0.5).half(), torch.randn(C, device=dev) * 0.3
    wf, gsum, bf = K.fold_layer_norm(w1, b1, gamma, beta)
    xs = x.float()
    stats = torch.stack([xs.sum(1), (xs * xs).sum(1)], 1).view(M, 1, 2).contiguous()
    # unfused: GEGLU GEMM (LayerNorm folded) + second GEMM with the residual
    w1a, b1a = K.pack_geglu(wf, bf, 128); _, g1a = K.pack_geglu(wf, gsum, 128)
    g = torch.empty(M, Fd, dtype=torch.float16, device=dev); out_a = torch.empty(M, C, dtype=torch.float16, device=dev)
    pa = K.Program()
    pa.add(K.linear_desc(x, w1a, g, bias=b1a, flags=L.EPI_GEGLU, bn=256, rownorm=(stats, g1a, 1e-5, C)))
    pa.add(K.linear_desc(g, K.pack_linear(w2), out_a, bias=b2, residual=x))
    w1p, b1p = K.pack_geglu(wf, bf, 64); _, g1p = K.pack_geglu(wf, gsum, 64)
    out_b = torch.empty_like(out_a)
    pb = K.Program(); pb.add(K.ff_fused_op(x, w1p, b1p, g1p, K.pack_linear(w2), b2, stats, out_b))
    pa.run(); pb.run(); torch.cuda.synchronize()
    err = float((out_a.double() - out_b.double()).norm() / out_a.double().norm())
    ta, tb = timeit(pa.run), timeit(pb.run)
    fl = 2.0 * M * C * 3 * Fd
    print(f"M={M}: GEGLU + FF2 {ta*1e3:7.1f} us ({fl/ta/1e9:6.0f} TF/s) | fused {tb*1e3:7.1f} us ({fl/tb/1e9:6.0f} TF/s) | rel diff {err:.1e}", flush=True)
