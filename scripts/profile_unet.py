"""Per-op timing table of one CFG UNet step (B = 2*b samples, 64x64 latent) and, with --eager-only,
a plain eager replay for ncu launch lists."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lightdiffusion_b200 import sd15
from lightdiffusion_b200.engine import UNetEngine, VAEDecoderEngine, VAEEncoderEngine

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--latent", type=int, default=64)
ap.add_argument("--eager-only", action="store_true")
ap.add_argument("--vae", action="store_true")
ap.add_argument("--vae-encode", action="store_true", help="--latent is the latent size; the image is 8x larger")
ap.add_argument("--out", default="gpurun_out/unet_ops.json")
args = ap.parse_args()
dev = torch.device("cuda", 0)
if args.vae_encode:
    sd = sd15.seeded_state_dict(sd15.vae_encoder_param_shapes(), 100, dtype=torch.float16, device=dev)
    eng = VAEEncoderEngine(sd, device=dev)
    plan = eng.plan(args.batch, args.latent * 8, args.latent * 8)
    plan.use_graph = False
    plan.run(torch.rand(args.batch, args.latent * 8, args.latent * 8, 3, device=dev),
             torch.randn(args.batch, 4, args.latent, args.latent, device=dev))
elif args.vae:
    sd = sd15.seeded_state_dict(sd15.vae_decoder_param_shapes(), 99, dtype=torch.float16, device=dev)
    eng = VAEDecoderEngine(sd, device=dev)
    plan = eng.plan(args.batch, args.latent, args.latent)
    plan.use_graph = False
    plan.run(torch.randn(args.batch, 4, args.latent, args.latent, device=dev))
else:
    sd = sd15.seeded_state_dict(sd15.unet_param_shapes(), 1234, dtype=torch.float16, device=dev)
    eng = UNetEngine(sd, device=dev)
    B = 2 * args.batch
    plan = eng.plan(B, args.latent, args.latent, cfg_pairs=True)
    plan.use_graph = False
    plan.set_context(torch.randn(B, 77, 768, device=dev))
    x = torch.randn(args.batch, 4, args.latent, args.latent, device=dev) * 10
    plan.run(x, torch.full((args.batch,), 5.0, device=dev), None, cfg_scale=7.0)
torch.cuda.synchronize()
if args.eager_only:
    plan.prog.run()
    torch.cuda.synchronize()
    sys.exit(0)
plan.prog.run_timed()
rows = plan.prog.run_timed()
tot = sum(r[3] for r in rows)
agg = {}
for kind, flops, tag, ms in rows:
    key = (kind, tag)
    a = agg.setdefault(key, [0, 0.0, 0.0])
    a[0] += 1; a[1] += ms; a[2] += flops
print(f"total {tot:.3f} ms over {len(rows)} ops")
out = []
for (kind, tag), (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    tf = fl / (ms * 1e-3) / 1e12 if ms > 0 else 0
    print(f"{kind:11s} {str(tag):28s} x{n:3d}  {ms:8.3f} ms  {100*ms/tot:5.1f}%  {tf:7.1f} TF/s(exec)")
    out.append({"kind": kind, "tag": tag, "n": n, "ms": ms, "tflops_exec": tf})
# the number that matters: CUDA-graph replay of the whole program (no events between ops, PDL overlap intact)
g = torch.cuda.CUDAGraph()
torch.cuda.synchronize()
with torch.cuda.graph(g):
    plan.prog.run()
for _ in range(3):
    g.replay()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    g.replay()
e1.record()
torch.cuda.synchronize()
graph_ms = e0.elapsed_time(e1) / 20
print(f"graph replay: {graph_ms:.3f} ms per program")
os.makedirs(os.path.dirname(args.out), exist_ok=True)
json.dump({"total_ms": tot, "graph_ms": graph_ms, "ops": out}, open(args.out, "w"), indent=1)
