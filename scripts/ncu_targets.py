"""One hot kernel per invocation, a few launches each, for `ncu --set full` captures (see scripts/gpu_ncu.sh).
Usage: python scripts/ncu_targets.py <case>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lightdiffusion_b200 import _lib as L, kernels as K

dev, case = "cuda", sys.argv[1]
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s, sc=1.0: (torch.randn(*s, device=dev, generator=g) * sc)
keep = []
if case == "gemm_dual_conv":          # 3x3 conv 1280 -> 1280 on 16 x 16 x 16: 64 pair tiles of 256 x 320, one wave
    x = rnd(16, 16, 16, 1280).half(); w = rnd(1280, 1280, 3, 3, sc=(9 * 1280) ** -0.5).half()
    wp, b = K.pack_conv3x3(w), rnd(1280); out = torch.empty(16, 16, 16, 1280, dtype=torch.float16, device=dev)
    op = K.conv3x3_desc(x, wp, out, bias=b); keep += [x, wp, b, out]
elif case == "gemm_dual_conv320":     # 3x3 conv 320 -> 320 on 16 x 64 x 64 (the r01 capture's shape): 256 x 320 tiles, 4 waves
    x = rnd(16, 64, 64, 320).half(); w = rnd(320, 320, 3, 3, sc=(9 * 320) ** -0.5).half()
    wp, b = K.pack_conv3x3(w), rnd(320); out = torch.empty(16, 64, 64, 320, dtype=torch.float16, device=dev)
    op = K.conv3x3_desc(x, wp, out, bias=b); keep += [x, wp, b, out]
elif case == "gemm_dual_lin":         # to_out + residual at the 16 x 16 level
    a = rnd(4096, 1280).half(); w = rnd(1280, 1280, sc=1280 ** -0.5).half(); b = rnd(1280); r = rnd(4096, 1280).half()
    out = torch.empty(4096, 1280, dtype=torch.float16, device=dev)
    op = K.linear_desc(a, w, out, bias=b, residual=r); keep += [a, w, b, r, out]
elif case == "conv_small_vae":        # VAE conv_out 128 -> 3 at 512^2, b = 8, fused process_output
    x = rnd(8, 512, 512, 128).half(); w = rnd(3, 128, 3, 3, sc=(9 * 128) ** -0.5).half(); b = rnd(3)
    wp = K.pack_conv3x3_small(w); o16 = torch.zeros(8, 512, 512, 8, dtype=torch.float16, device=dev)
    img = torch.empty(8, 512, 512, 3, dtype=torch.float32, device=dev)
    op = K.conv3x3_small_op(x, wp, b, 3, out16=o16, out_img=img, post=1); keep += [x, wp, b, o16, img]
elif case in ("halo_conv320", "halo_conv1280", "halo_vae128"):   # GroupNorm + SiLU + 3x3 conv through the halo-tile kernel
    B, H, C = {"halo_conv320": (16, 64, 320), "halo_conv1280": (16, 16, 1280), "halo_vae128": (8, 512, 128)}[case]
    x = (rnd(B, H, H, C) * 1.3 + 0.3).half(); w = rnd(C, C, 3, 3, sc=(9 * C) ** -0.5).half()
    wp, b = K.pack_conv3x3(w), rnd(C); out = torch.empty(B, H, H, C, dtype=torch.float16, device=dev)
    gam, bet = rnd(C) * 0.1 + 1, rnd(C) * 0.1
    coef = torch.empty(B, C, 2, dtype=torch.float32, device=dev)
    st = torch.empty(B * L.GN_STATS_FLOATS_PER_BATCH, dtype=torch.float32, device=dev)
    tk = torch.zeros(B, dtype=torch.int32, device=dev)
    K.run_op(K.group_norm_coef_op(x, None, gam, bet, coef, st, tk, 1e-5))
    op = K.conv3x3_halo_desc([x], wp, out, gn=(coef, True), bias=b); keep += [x, wp, b, out, gam, bet, coef, st, tk]
elif case == "xattn40":               # cross-attention over the 77 context tokens at the 64 x 64 level
    q = rnd(16, 4096, 320).half(); k, v = rnd(16, 77, 320).half(), rnd(16, 77, 320).half(); o = torch.empty_like(q)
    op = K.attention_op(q, k, v, o, 8); keep += [q, k, v, o]
elif case == "attention_d40":
    q, k, v = (rnd(16, 4096, 320).half() for _ in range(3)); o = torch.empty_like(q)
    op = K.attention_op(q, k, v, o, 8); keep += [q, k, v, o]
elif case == "gn_colstats_vae":       # GroupNorm + SiLU of a 8 x 512 x 512 x 128 map with producer statistics
    xin = rnd(8, 512, 512, 128).half(); w = rnd(128, 128, 3, 3, sc=(9 * 128) ** -0.5).half(); wp, b = K.pack_conv3x3(w), rnd(128)
    x = torch.empty(8, 512, 512, 128, dtype=torch.float16, device=dev)
    d = K.conv3x3_desc(xin, wp, x, bias=b); cs = K.attach_colstats(d, 8, 128); K.run_gemm(d)
    gam, bet = rnd(128) * 0.1 + 1, rnd(128) * 0.1; y = torch.empty_like(x)
    st = torch.empty(8 * L.GN_STATS_FLOATS_PER_BATCH, dtype=torch.float32, device=dev)
    op = K.group_norm_op(x, None, gam, bet, y, st, 1e-6, True, colstats=(cs, None)); keep += [xin, wp, b, x, cs, gam, bet, y, st]
elif case == "ff_fused":              # x + ff(norm3(x)) at the 64 x 64 level in one launch (C = 320, 16 samples)
    M, C, Fd = 65536, 320, 1280
    x = (rnd(M, C) * 1.2 + 0.3).half(); gam, bet = rnd(C) * 0.1 + 1, rnd(C) * 0.1
    w1, b1 = rnd(2 * Fd, C, sc=C ** -0.5).half(), rnd(2 * Fd) * 0.3
    w2, b2 = rnd(C, Fd, sc=Fd ** -0.5).half(), rnd(C) * 0.3
    wf, gsum, bf = K.fold_layer_norm(w1, b1, gam, bet)
    w1p, b1p = K.pack_geglu(wf, bf, 64); _, g1p = K.pack_geglu(wf, gsum, 64)
    xs = x.float(); stats = torch.stack([xs.sum(1), (xs * xs).sum(1)], 1).view(M, 1, 2).contiguous()
    out = torch.empty(M, C, dtype=torch.float16, device=dev); w2p = K.pack_linear(w2)
    op = K.ff_fused_op(x, w1p, b1p, g1p, w2p, b2, stats, out); keep += [x, w1p, b1p, g1p, w2p, b2, stats, out]
else:
    raise SystemExit("unknown case " + case)
prog = K.Program(); prog.add(op)
for _ in range(4):
    prog.run()
torch.cuda.synchronize()
print("ok", case)
