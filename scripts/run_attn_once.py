import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lightdiffusion_b200 import kernels as K
B, N, H, d = 16, 4096, 8, 40
q = torch.randn(B, N, H * d, device="cuda").half(); k = torch.randn_like(q); v = torch.randn_like(q); o = torch.empty_like(q)
prog = K.Program(); prog.add(K.attention_op(q, k, v, o, H))
for _ in range(3): prog.run()
torch.cuda.synchronize()
