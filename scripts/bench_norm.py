"""Steady-state time of GroupNorm / LayerNorm shapes of the SD1.5 UNet: 20 back-to-back launches on rotating
buffers inside one recorded program (no host work between launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lightdiffusion_b200 import _lib as L, kernels as K
dev = "cuda"
def timeit(prog, iters=5):
    prog.run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): prog.run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
REP = 20
for (B, HW, C, silu) in [(16, 4096, 320, True), (16, 1024, 640, True), (16, 256, 1280, True), (16, 64, 1280, True), (16, 4096, 640, True), (16, 4096, 960, True), (16, 1024, 1920, True)]:
    hh = int(HW ** 0.5)
    xs = [torch.randn(B, hh, HW // hh, C, device=dev).half() for _ in range(3)]
    ys = [torch.empty_like(xs[0]) for _ in range(3)]
    g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
    stats = torch.zeros(B * L.GN_STATS_FLOATS_PER_BATCH, dtype=torch.float32, device=dev)
    prog = K.Program()
    for i in range(REP):
        prog.add(K.group_norm_op(xs[i % 3], None, g, b, ys[i % 3], stats, 1e-5, silu))
    ms = timeit(prog) / REP
    byts = B * HW * C * 2 * 2
    print(f"group_norm B={B} HW={HW} C={C}: {ms*1e3:7.1f} us   {byts/ms/1e9:6.2f} TB/s (r+w algorithmic)", flush=True)
for (rows, C) in [(65536, 320), (16384, 640), (4096, 1280)]:
    xs = [torch.randn(rows, C, device=dev).half() for _ in range(3)]
    ys = [torch.empty_like(xs[0]) for _ in range(3)]
    g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
    prog = K.Program()
    for i in range(REP):
        prog.add(K.layer_norm_op(xs[i % 3], g, b, ys[i % 3]))
    ms = timeit(prog) / REP
    byts = rows * C * 2 * 2
    print(f"layer_norm rows={rows} C={C}: {ms*1e3:7.1f} us   {byts/ms/1e9:6.2f} TB/s (r+w algorithmic)", flush=True)
