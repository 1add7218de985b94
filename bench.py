#!/usr/bin/env python
"""Benchmark of the SD1.5 denoising hot path (BASELINE.json metric):

    SD1.5 512x512 images/sec, 20-step DPM++ 2M Karras, CFG 7, fp16, random-init weights, synthetic inputs.

One "step" = one pass of the hot path over one batch: 20 sampler steps (each one CFG UNet evaluation of
2*b samples + the fused DPM++ 2M update) followed by the VAE decode of the b latents, and -- with more
than one GPU -- the final gather of the decoded images.  Images are sharded over ranks (weak scaling:
b images per GPU); there is no collective on the sampler-step path.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference algorithm's CPU path on the host cores

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "sd15_512x512_images_per_sec_20step_dpmpp2m_cfg"
UNIT = "images/s"
STEPS_PER_IMAGE = 20
CFG_SCALE = 7.0
# Algorithmic FLOPs (SURVEY.md §8d / BASELINE.md §2; counted by hooks on the reference modules)
GF_UNET_SAMPLE = 803.3          # one UNet evaluation of one sample at a 64x64 latent
GF_UNET_GEMM_SAMPLE = 677.2     # of which conv3x3 400.3 + conv1x1 43.6 + linear 233.3 (the implicit-GEMM kernel)
GF_UNET_ATTN_SAMPLE = 126.1
GF_VAE_IMAGE = 2514.5
TF_IMAGE = 34.65                # 20 steps x 2 (CFG) x UNet + VAE decode


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"tflops_burst": d.get("bf16_tflops"), "tflops_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d.get("hbm_gbs"), "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops_burst": 1590.0, "tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


# ----------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], None, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [t.strip() for t in ln.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[1])); smax = float(p[2]); power.append(float(p[3]))
            except ValueError:
                continue
            for n, v in zip(names, p[4:8]):
                if v == "Active":
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": smax, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": smax, "reasons": sorted(reasons),
                "power_w_max": max(power) if power else None, "samples": len(sm)}


# ----------------------------------------------------------------------------------- CPU legs (oracle)
def cpu_reference_sample(threads: int):
    """Times the reference algorithm (oracle port, fp32) on the host cores on a bounded sample of the
    workload: ONE CFG UNet evaluation for one 512^2 image (2 samples, 64x64 latent) and ONE VAE decode of
    a 16x16 latent patch (1/16 of the image; conv work is linear in area).  images/s is then
    1 / (20 * t_unet + 16 * t_vae).  Returns (images_per_s, detail)."""
    from lightdiffusion_b200 import sd15
    from oracle import ldm_oracle as O
    torch.set_num_threads(threads)
    cfg = sd15.SD15_UNET
    usd = sd15.seeded_state_dict(sd15.unet_param_shapes(cfg), 1234)
    vsd = sd15.seeded_state_dict(sd15.vae_decoder_param_shapes(), 99)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 4, 64, 64, generator=g) * 14.6
    cond, uncond = torch.randn(1, 77, 768, generator=g), torch.randn(1, 77, 768, generator=g)
    model = O.cfg_denoiser(usd, cfg, cond, uncond, CFG_SCALE)
    z = torch.randn(1, 4, 16, 16, generator=g)
    with torch.inference_mode():
        t0 = time.perf_counter()
        model(x, torch.tensor([14.6146]))
        t_unet = time.perf_counter() - t0
        t0 = time.perf_counter()
        O.vae_decode(vsd, sd15.SD15_VAE, z)
        t_vae = time.perf_counter() - t0
    per_image = STEPS_PER_IMAGE * t_unet + 16 * t_vae
    return 1.0 / per_image, {"t_unet_cfg_step_s": round(t_unet, 3), "t_vae_16x16_patch_s": round(t_vae, 3)}


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    vals, detail = [], {}
    for i in range(args.warmup + args.steps):
        v, detail = cpu_reference_sample(cores)
        if i >= args.warmup:
            vals.append(v)
        if i == 0 and args.warmup > 1:        # keep the whole run within minutes: one warm-up sample is enough
            args.warmup = 1
    value = 1.0 / statistics.mean(1.0 / v for v in vals)
    sample = ("per step: 1 CFG UNet eval (2 samples, 64x64 latent, fp32) + VAE decode of a 16x16 latent patch; "
              "images/s = 1/(20*t_unet + 16*t_vae)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(vals), "warmup": args.warmup, "ms_per_step": 1e3 / value, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args, "cpu"),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, **detail},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def bench_config(args, where):
    return {"workload": f"SD1.5 txt2img 512x512, batch {args.batch} per GPU, fp16, 20 steps DPM++ 2M Karras, CFG 7"
                        + (", images sharded over GPUs" if args.gpus > 1 else ", on 1 B200"),
            "images_per_gpu": args.batch, "global_batch": args.batch * args.gpus, "latent": "64x64",
            "sampler_steps": STEPS_PER_IMAGE, "cfg_scale": CFG_SCALE, "scheduler": "karras",
            "weights": "random-init SD1.5 architecture (seeded)", "parallelism": f"images sharded dp{args.gpus}",
            "l2": "working set (1.7 GB fp16 weights + activations) exceeds the 126 MB L2; no flush needed",
            "where": where}


# ----------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch.distributed as dist
    from lightdiffusion_b200 import sd15
    from lightdiffusion_b200.engine import UNetEngine, VAEDecoderEngine
    from lightdiffusion_b200.pipeline import Txt2ImgPipeline, calculate_sigmas, gather_images

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    b = args.batch
    peaks = measured_peaks()

    usd = sd15.seeded_state_dict(sd15.unet_param_shapes(), 1234, dtype=torch.float16, device=dev)
    vsd = sd15.seeded_state_dict(sd15.vae_decoder_param_shapes(), 99, dtype=torch.float16, device=dev)
    pipe = Txt2ImgPipeline(UNetEngine(usd, device=dev), VAEDecoderEngine(vsd, device=dev))
    del usd, vsd
    sig = calculate_sigmas("karras", STEPS_PER_IMAGE)

    # synthetic inputs, seeded per global image index (results do not depend on the sharding)
    def host_inputs():
        noise = torch.stack([torch.randn(4, 64, 64, generator=torch.Generator().manual_seed(42 + rank * b + i))
                             for i in range(b)])
        g = torch.Generator().manual_seed(7 + rank)
        return noise, torch.randn(b, 77, 768, generator=g), torch.randn(b, 77, 768, generator=g)
    h_noise, h_cond, h_uncond = (t.pin_memory() for t in host_inputs())
    d_noise, d_cond, d_uncond = h_noise.to(dev), h_cond.to(dev), h_uncond.to(dev)
    h_img = torch.empty(b, 512, 512, 3, dtype=torch.float32).pin_memory()

    def step_resident():
        z, img = pipe.generate(d_noise, d_cond, d_uncond, CFG_SCALE, sig, "dpmpp_2m")
        if world > 1:
            gather_images(img, b * world)
        return img

    def step_e2e():
        n, c, u = h_noise.to(dev, non_blocking=True), h_cond.to(dev, non_blocking=True), h_uncond.to(dev, non_blocking=True)
        z, img = pipe.generate(n, c, u, CFG_SCALE, sig, "dpmpp_2m")
        if world > 1:
            gather_images(img, b * world)
        h_img.copy_(img, non_blocking=True)

    def timed(fn, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed(step_resident, args.steps)
    clocks = sampler.stop()
    launches_per_step = pipe.gpu_launches
    value = world * b * args.steps / (ms * 1e-3)

    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    e2e_value = world * b * args.steps / (ms_e2e * 1e-3)
    h2d = sum(t.numel() * t.element_size() for t in (h_noise, h_cond, h_uncond))
    d2h = h_img.numel() * h_img.element_size()

    # UNet step time (graph replay of the 2b-sample CFG evaluation) and per-kernel roofline
    plan = pipe.unet.plan(2 * b, 64, 64, cfg_pairs=True)
    for _ in range(3):
        plan.run(plan.x_in, plan.sigma_in, None, cfg_scale=CFG_SCALE)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(STEPS_PER_IMAGE):
        plan.run(plan.x_in, plan.sigma_in, None, cfg_scale=CFG_SCALE)
    e1.record()
    torch.cuda.synchronize()
    unet_ms = e0.elapsed_time(e1) / STEPS_PER_IMAGE
    plan.prog.run_timed()
    per_op = plan.prog.run_timed()                   # CUDA events between ops, on the launching stream
    t_kind, n_kind = {}, {}
    for kind, flops, tag, t in per_op:
        t_kind[kind] = t_kind.get(kind, 0.0) + t
        n_kind[kind] = n_kind.get(kind, 0) + 1
    t_total = sum(t_kind.values())
    gemm_ms = t_kind.get("gemm", 0.0)
    achieved = GF_UNET_GEMM_SAMPLE * 1e9 * 2 * b / (gemm_ms * 1e-3) / 1e12
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")     # committed ncu measurement (per launch)
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic, traffic_src = tj.get("dram_bytes_per_launch"), "profiles/r01_gemm_traffic.json: " + tj.get("source", "")
    roofline = {"kernel": "ldx_gemm_kernel (implicit-GEMM conv / linear, tcgen05)", "bound": "tensor",
                "achieved": achieved, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["tflops_sustained"], "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peaks["source"] + ", sustained bf16 cuBLAS figure (kernel timed inside a long step)",
                "launches_per_unet_step": n_kind.get("gemm", 0), "ms_per_unet_step": gemm_ms,
                "share_of_unet_step": gemm_ms / t_total,
                "algorithmic_gflop_per_launch_avg": GF_UNET_GEMM_SAMPLE * 2 * b / max(1, n_kind.get("gemm", 0)),
                "frac_of_nominal_2250": achieved / 2250.0}
    breakdown = {k: {"ms": round(v, 4), "launches": n_kind[k], "share": round(v / t_total, 4)} for k, v in t_kind.items()}
    if "attention" in t_kind:
        breakdown["attention"]["tflops"] = round(GF_UNET_ATTN_SAMPLE * 1e9 * 2 * b / (t_kind["attention"] * 1e-3) / 1e12, 1)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, detail = cpu_reference_sample(cores)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "1 CFG UNet eval (2 samples, 64x64 latent, fp32) + VAE decode of a 16x16 latent patch; "
                                  "images/s = 1/(20*t_unet + 16*t_vae)", **detail}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
                "config": bench_config(args, "gpu"),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches_per_step * args.steps,
                "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
                "unet_step_ms": unet_ms, "unet_step_tflops": GF_UNET_SAMPLE * 1e9 * 2 * b / (unet_ms * 1e-3) / 1e12,
                "pipeline_tflops_per_gpu": TF_IMAGE * value / world,
                "frac_of_sustained_peak": TF_IMAGE * value / world / peaks["tflops_sustained"],
                "unet_breakdown": breakdown}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="images per GPU (BASELINE.json config 2: 8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
