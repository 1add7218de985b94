"""Phase cycle counts of the short-context cross-attention kernel (variant build with -DLDX_XA_TIMING)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lightdiffusion_b200 import _lib as L, kernels as K
B, Nq, Nk, H, d = 16, 4096, 77, 8, int(sys.argv[1]) if len(sys.argv) > 1 else 40
if d == 80: Nq = 1024
q = torch.randn(B, Nq, H * d, device="cuda").half(); k = torch.randn(B, Nk, H * d, device="cuda").half(); v = torch.randn(B, Nk, H * d, device="cuda").half()
o = torch.empty_like(q)
prog = K.Program(); prog.add(K.attention_op(q, k, v, o, H))
for _ in range(5): prog.run()
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 32)()
fn = L.lib().ldx_debug_xattn_ts; fn.restype = ctypes.c_int; fn.argtypes = [ctypes.c_void_p]
assert fn(buf) == 0
n = buf[8]
names = ["wait s_full", "S ld + release", "mask/max/exp/P store", "fence + arrive p_full", "wait o_full", "O readout + store", "-", "loop / other"]
print(f"softmax warp 4 of CTA 0: {n} units in the CTA ({(n + 1) // 2} for this warpgroup), cycles per own unit:")
for i, nm in enumerate(names):
    print(f"   {nm:24s} {buf[i] / ((n + 1) // 2):9.0f}")
print(f"   total                    {sum(buf[i] for i in range(8)) / ((n + 1) // 2):9.0f}")
names = ["wait qk_full", "wait s_free", "issue S", "wait p_full", "wait v_full", "issue PV", "-", "loop / other"]
n = buf[24]
print(f"MMA warp of CTA 0: {n} units, cycles per unit:")
for i, nm in enumerate(names):
    print(f"   {nm:24s} {buf[16 + i] / max(n, 1):9.0f}")
print(f"   total                    {sum(buf[16 + i] for i in range(8)) / max(n, 1):9.0f}")
