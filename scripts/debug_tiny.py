import torch, torch.nn.functional as F, math, sys
sys.path.insert(0, '.')
from lightdiffusion_b200 import _lib as L, kernels as K
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
DEV='cuda'
def rel(a,b): a,b=a.double(),b.double(); return float((a-b).norm()/(b.norm()+1e-30))
def rnd(*s, seed=0, scale=1.0):
    g=torch.Generator().manual_seed(seed); return (torch.randn(*s,generator=g)*scale).to(DEV)
h16=lambda t:t.half()
for (B,H,W,C) in [(1,1,1,256),(1,2,2,256),(2,1,1,256),(1,4,4,256),(3,2,2,128)]:
    x=h16(rnd(B,H,W,C,seed=1)); w=h16(rnd(C,C,3,3,seed=2,scale=(9*C)**-0.5)); bias=rnd(C,seed=3)
    out=torch.empty(B,H,W,C,dtype=torch.float16,device=DEV)
    K.run_gemm(K.conv3x3_desc(x,K.pack_conv3x3(w),out,bias=bias))
    ref=F.conv2d(x.float().permute(0,3,1,2),w.float(),bias,padding=1)
    print('conv3x3',(B,H,W,C),rel(out.permute(0,3,1,2),ref))
    if H%2==0:
        out=torch.empty(B,H//2,W//2,C,dtype=torch.float16,device=DEV)
        K.run_gemm(K.conv3x3_s2_desc(x,K.pack_conv3x3(w),out,bias=bias))
        ref=F.conv2d(x.float().permute(0,3,1,2),w.float(),bias,stride=2,padding=1)
        print('  s2',rel(out.permute(0,3,1,2),ref))
    out=torch.empty(B,2*H,2*W,C,dtype=torch.float16,device=DEV)
    for d in K.upconv3x3_descs(x,K.pack_upconv3x3(w),out,bias=bias): K.run_gemm(d)
    up=F.interpolate(x.float().permute(0,3,1,2),scale_factor=2.0,mode='nearest')
    print('  upconv',rel(out.permute(0,3,1,2),F.conv2d(up,w.float(),bias,padding=1)))
    gamma,beta=rnd(C,seed=4)*0.2+1,rnd(C,seed=5)*0.2
    y=torch.empty(B,H,W,C,dtype=torch.float16,device=DEV); st=torch.empty(B*L.GN_STATS_FLOATS_PER_BATCH,device=DEV)
    K.run_op(K.group_norm_op(x,None,gamma,beta,y,st,1e-5,True))
    print('  gn',rel(y.permute(0,3,1,2),F.silu(F.group_norm(x.float().permute(0,3,1,2),32,gamma,beta,1e-5))))
    heads,d=2,C//2
    q=h16(rnd(B,H*W,C,seed=6)); k=h16(rnd(B,H*W,C,seed=7)); v=h16(rnd(B,H*W,C,seed=8)); o=torch.empty_like(q)
    K.run_op(K.attention_op(q,k,v,o,heads))
    qf,kf,vf=(t.float().view(B,-1,heads,d).transpose(1,2) for t in (q,k,v))
    print('  attn',rel(o,F.scaled_dot_product_attention(qf,kf,vf).transpose(1,2).reshape(B,H*W,C)))
    a=h16(rnd(B*H*W,C,seed=9)); wl=h16(rnd(C,C,seed=10,scale=C**-0.5)); o2=torch.empty(B*H*W,C,dtype=torch.float16,device=DEV)
    K.run_gemm(K.linear_desc(a,wl,o2)); print('  lin',rel(o2,a.float()@wl.float().t()))
    yl=torch.empty_like(a); K.run_op(K.layer_norm_op(a,gamma,beta,yl)); print('  ln',rel(yl,F.layer_norm(a.float(),(C,),gamma,beta,1e-5)))
