import torch, torch.nn.functional as F
dev="cuda"
def t(fn, iters=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for (B,C,H) in [(16,320,64),(16,640,32),(16,1280,16)]:
    xs=[torch.randn(B,C,H,H,device=dev).half() for _ in range(3)]
    g=torch.ones(C,device=dev).half(); b=torch.zeros(C,device=dev).half()
    i=[0]
    def f():
        F.silu(F.group_norm(xs[i[0]%3],32,g,b,1e-5)); i[0]+=1
    def f2():
        F.group_norm(xs[i[0]%3],32,g,b,1e-5); i[0]+=1
    xl=[x.to(memory_format=torch.channels_last) for x in xs]
    def f3():
        F.group_norm(xl[i[0]%3],32,g,b,1e-5); i[0]+=1
    print(f"torch GN+SiLU NCHW {B}x{C}x{H}x{H}: {t(f)*1e3:.1f} us; GN only: {t(f2)*1e3:.1f} us; GN channels_last: {t(f3)*1e3:.1f} us")
