"""Variant build only (-DLDX_GEMM_TIMING): prints CTA 0's pipeline milestones of one GEMM launch."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lightdiffusion_b200 import _lib as L, kernels as K
dev = "cuda"
lib = L.lib()
lib.ldx_debug_gemm_ts.argtypes = [C.POINTER(C.c_ulonglong)]
names = ["kernel entry", "prologue done (barriers, TMEM)", "griddepcontrol.wait done", "first stage landed", "last stage of tile 0 landed",
         "accumulator 0 complete", "epilogue loop done", "stores drained"]
for (M, N, Kd, res) in [(65536, 320, 320, True), (4096, 1280, 1280, True), (16384, 640, 640, True), (4096, 1280, 5120, False), (65536, 2560, 320, 'geglu'), (65536, 320, 1280, True)]:
    a = torch.randn(M, Kd, device=dev).half(); w = (torch.randn(N, Kd, device=dev) * Kd ** -0.5).half()
    b = torch.randn(N, device=dev); r = torch.randn(M, N, device=dev).half() if res is True else None
    out = torch.empty(M, N, dtype=torch.float16, device=dev)
    if res == 'geglu':
        w, b = K.pack_geglu(w, b, 128)
        out = torch.empty(M, N // 2, dtype=torch.float16, device=dev)
        d = K.linear_desc(a, w, out, bias=b, flags=L.EPI_GEGLU, bn=256)
    else:
        d = K.linear_desc(a, w, out, bias=b, residual=r)
    prog = K.Program(); prog.add(d)
    for _ in range(3):
        prog.run()
    torch.cuda.synchronize()
    ts = (C.c_ulonglong * 16)()
    lib.ldx_debug_gemm_ts(ts)
    t0 = ts[0]
    print(f"== M={M} N={N} K={Kd} res={res}")
    for i, n in enumerate(names):
        print(f"   {n:36s} +{(ts[i] - t0) / 1e3:8.2f} us")
    nch = max(1, ts[15])
    ph = ["wait accumulator", "slot drain wait + TMA store issue", "bias + math (after TMEM wait)", "TMEM load + wait", "pack + STS", "other", "fence.proxy.async"]
    print(f"   warp 2 of CTA 0: {ts[15]} chunks; cycles per chunk by phase:")
    for i, n in enumerate(ph):
        print(f"      {n:40s} {ts[8 + i] / nch:8.0f}")
