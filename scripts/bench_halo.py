"""GroupNorm + SiLU + 3x3 conv on the ResBlock shapes of the SD1.5 UNet / VAE: the two-pass GroupNorm followed by the
implicit-GEMM conv against the statistics-only GroupNorm followed by the halo-tile conv with the normalisation in
its operand path.  Steady-state time per launch sequence (20 back-to-back replays, CUDA events), and with --cold the
L2 flushed before each round."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lightdiffusion_b200 import _lib as L, kernels as K

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=7)
ap.add_argument("--cold", action="store_true")
ap.add_argument("--case", default="all")
args = ap.parse_args()
dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn):
    fn(); torch.cuda.synchronize()
    ts, reps = [], (1 if args.cold else 20)
    for _ in range(args.iters):
        if args.cold:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps)
    ts.sort()
    return ts[len(ts) // 2]


def case(B, H, W, Cin, Cout, c1=0):
    x0 = (torch.randn(B, H, W, Cin, device=dev) * 1.3 + 0.3).half()
    x1 = torch.randn(B, H, W, c1, device=dev).half() if c1 else None
    C = Cin + c1
    w = (torch.randn(Cout, C, 3, 3, device=dev) * (9 * C) ** -0.5).half()
    wp = K.pack_conv3x3(w)
    bias, temb = torch.randn(Cout, device=dev), torch.randn(B, Cout, device=dev)
    gamma, beta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.2
    stats = torch.empty(B * L.GN_STATS_FLOATS_PER_BATCH, dtype=torch.float32, device=dev)
    y = torch.empty(B, H, W, C, dtype=torch.float16, device=dev)
    out_a = torch.empty(B, H, W, Cout, dtype=torch.float16, device=dev)
    out_b = torch.empty_like(out_a)
    out_c = torch.empty_like(out_a)
    coef = torch.empty(B, C, 2, dtype=torch.float32, device=dev)
    tickets = torch.zeros(B, dtype=torch.int32, device=dev)
    flops = 2.0 * B * H * W * Cout * 9 * C

    pa = K.Program()                         # current path
    pa.add(K.group_norm_op(x0, x1, gamma, beta, y, stats, 1e-5, True))
    pa.add(K.conv3x3_desc(y, wp, out_a, bias=bias, rowadd=temb))
    pa_conv = K.Program()
    pa_conv.add(K.conv3x3_desc(y, wp, out_a, bias=bias, rowadd=temb))
    pb = K.Program()                         # halo conv with the normalisation in its operand path
    pb.add(K.group_norm_coef_op(x0, x1, gamma, beta, coef, stats, tickets, 1e-5))
    srcs = [x0] + ([x1] if x1 is not None else [])
    pb.add(K.conv3x3_halo_desc(srcs, wp, out_b, gn=(coef, True), bias=bias, rowadd=temb))
    pb_conv = K.Program()
    pb_conv.add(K.conv3x3_halo_desc(srcs, wp, out_b, gn=(coef, True), bias=bias, rowadd=temb))
    pc_conv = K.Program()                    # halo conv alone, no transform (operand traffic effect only)
    pc_conv.add(K.conv3x3_halo_desc([y], wp, out_c, bias=bias, rowadd=temb))
    pa.run(); pb.run(); pc_conv.run(); torch.cuda.synchronize()
    err = float((out_a.double() - out_b.double()).norm() / out_a.double().norm())
    err_c = float((out_a.double() - out_c.double()).norm() / out_a.double().norm())
    ta, tac, tb, tbc, tcc = timeit(pa.run), timeit(pa_conv.run), timeit(pb.run), timeit(pb_conv.run), timeit(pc_conv.run)
    tf = lambda t: flops / (t * 1e-3) / 1e12
    print(f"B{B} {H}x{W} {C}->{Cout}: GN+conv {ta*1e3:7.1f} us (conv {tac*1e3:6.1f} us {tf(tac):6.0f} TF) | coef+halo "
          f"{tb*1e3:7.1f} us (halo+gn {tbc*1e3:6.1f} us {tf(tbc):6.0f} TF) | halo plain {tcc*1e3:6.1f} us {tf(tcc):6.0f} TF | "
          f"rel diff {err:.1e} / {err_c:.1e}", flush=True)


CASES = {
    "u320": (16, 64, 64, 320, 320, 0),
    "u640": (16, 32, 32, 640, 640, 0),
    "u1280": (16, 16, 16, 1280, 1280, 0),
    "u320_640": (16, 32, 32, 320, 640, 0),
    "u640_1280": (16, 16, 16, 640, 1280, 0),
    "up2560": (16, 16, 16, 1280, 1280, 1280),
    "up1920": (16, 16, 16, 1280, 1280, 640),
    "up1280_32": (16, 32, 32, 640, 640, 640),
    "up960_32": (16, 32, 32, 640, 640, 320),
    "up960": (16, 64, 64, 640, 320, 320),
    "up640": (16, 64, 64, 320, 320, 320),
    "vae512": (8, 64, 64, 512, 512, 0),
    "vae512_128": (8, 128, 128, 512, 512, 0),
    "vae256": (8, 256, 256, 256, 256, 0),
    "vae128": (8, 512, 512, 128, 128, 0),
}
for n in (list(CASES) if args.case == "all" else args.case.split(",")):
    case(*CASES[n])
