"""Times single GEMM / conv shapes through the C ABI (CUDA events, L2 flushed between iterations)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lightdiffusion_b200 import _lib as L, kernels as K

ap = argparse.ArgumentParser()
ap.add_argument("--case", default="all")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--hot", action="store_true", help="no L2 flush: inputs stay L2-resident between iterations")
args = ap.parse_args()
dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

def timeit(fn, iters):
    """Median over `iters` rounds of the per-launch time of 20 back-to-back replays (steady state, as inside
    the UNet program); with --flush the L2 is flushed before every round instead (cold first launch)."""
    fn(); torch.cuda.synchronize()
    ts = []
    reps = 1 if not args.hot else 20
    for _ in range(iters):
        if not args.hot:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps)
    ts.sort()
    return ts[len(ts) // 2]

def prog_of(d):
    """A recorded one-op program: prepared once (tensor maps, split-K workspace), replayed per iteration."""
    prog = K.Program()
    prog.add(d)
    return prog.run

def lin(M, N, Kd, res=False, bias=True, geglu=False, bn=0):
    a = torch.randn(M, Kd, device=dev).half(); w = (torch.randn(N, Kd, device=dev) * Kd ** -0.5).half()
    b = torch.randn(N, device=dev) if bias else None
    r = torch.randn(M, N, device=dev).half() if res else None
    if geglu:
        w, b = K.pack_geglu(w, b, 128)
        out = torch.empty(M, N // 2, dtype=torch.float16, device=dev)
        d = K.linear_desc(a, w, out, bias=b, flags=L.EPI_GEGLU, bn=256)
    else:
        out = torch.empty(M, N, dtype=torch.float16, device=dev)
        d = K.linear_desc(a, w, out, bias=b, residual=r, bn=bn)
    return prog_of(d), 2.0 * M * N * Kd, (a, w, b, r, out)

def conv(B, H, W, Cin, Cout, res=False):
    x = torch.randn(B, H, W, Cin, device=dev).half(); w = (torch.randn(Cout, Cin, 3, 3, device=dev) * (9 * Cin) ** -0.5).half()
    wp = K.pack_conv3x3(w); b = torch.randn(Cout, device=dev)
    r = torch.randn(B, H, W, Cout, device=dev).half() if res else None
    out = torch.empty(B, H, W, Cout, dtype=torch.float16, device=dev)
    d = K.conv3x3_desc(x, wp, out, bias=b, residual=r)
    return prog_of(d), 2.0 * B * H * W * Cout * 9 * Cin, (x, wp, b, r, out)

CASES = {
    "lin320": lambda: lin(65536, 320, 320),
    "lin320_res": lambda: lin(65536, 320, 320, res=True),
    "lin640_res": lambda: lin(16384, 640, 640, res=True),
    "lin1280_res": lambda: lin(4096, 1280, 1280, res=True),
    "geglu320": lambda: lin(65536, 2560, 320, geglu=True),
    "geglu1280": lambda: lin(4096, 10240, 1280, geglu=True),
    "ff2_320": lambda: lin(65536, 320, 1280, res=True),
    "qkv320": lambda: lin(65536, 960, 320, bias=False),
    "conv320": lambda: conv(16, 64, 64, 320, 320),
    "conv320_res": lambda: conv(16, 64, 64, 320, 320, res=True),
    "conv640": lambda: conv(16, 32, 32, 640, 640),
    "conv1280": lambda: conv(16, 16, 16, 1280, 1280),
    "conv1280_8": lambda: conv(16, 8, 8, 1280, 1280),
    "conv2560_8": lambda: conv(16, 8, 8, 2560, 1280),
}
names = list(CASES) if args.case == "all" else args.case.split(",")
for n in list(names):
    if n.startswith("lin:"):            # lin:M:N:K[:res]
        f = n.split(":")
        M_, N_, K_ = int(f[1]), int(f[2]), int(f[3])
        CASES[n] = (lambda M_=M_, N_=N_, K_=K_, r_=len(f) > 4: lin(M_, N_, K_, res=r_))
for n in names:
    fn, fl, keep = CASES[n]()
    ms = timeit(fn, args.iters)
    print(f"{n:14s} {ms*1e3:9.1f} us  {fl/ms/1e9:8.1f} TF/s", flush=True)
