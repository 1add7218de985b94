"""Phase cycle counts of the first softmax warp of the flash self-attention kernel (variant build -DLDX_ATT_TIMING)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lightdiffusion_b200 import _lib as L, kernels as K
B, N, H, d = 16, 4096, 8, int(sys.argv[1]) if len(sys.argv) > 1 else 40
if d == 80: N = 1024
q, k, v = (torch.randn(B, N, H * d, device="cuda").half() for _ in range(3))
o = torch.empty_like(q)
prog = K.Program(); prog.add(K.attention_op(q, k, v, o, H))
for _ in range(3): prog.run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): prog.run()
e1.record(); torch.cuda.synchronize()
print(f"d={d} N={N}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per launch")
buf = (ctypes.c_ulonglong * 16)()
fn = L.lib().ldx_debug_att_ts; fn.restype = ctypes.c_int; fn.argtypes = [ctypes.c_void_p]
assert fn(buf) == 0
n = buf[8]
names = ["wait s_full", "S ld + release", "mask / max / rescale test", "wait pv_done (P buffer)", "exp + pack + P store", "fence + arrive p_full", "wait MUFU token", "loop / other"]
print(f"first softmax warp of CTA 0: {n} K/V tiles, cycles per tile:")
for i, nm in enumerate(names):
    print(f"   {nm:28s} {buf[i] / max(n, 1):9.0f}")
print(f"   total                        {sum(buf[i] for i in range(8)) / max(n, 1):9.0f}")
