#!/usr/bin/env python
"""Benchmark of the SD1.5 denoising hot path (BASELINE.json metric):

    SD1.5 512x512 images/sec, 20-step DPM++ 2M Karras, CFG 7, fp16, random-init weights, synthetic inputs.

One "step" = one pass of the hot path over one batch: 20 sampler steps (each one CFG UNet evaluation of
2*b samples + the fused CFG / DPM++ 2M tail kernel) followed by the VAE decode of the b latents, and -- with more
than one GPU -- the final gather of the decoded images to rank 0.  Images are sharded over ranks (weak scaling:
b images per GPU; ``--total-images N`` fixes the total instead, BASELINE config 3); there is no collective on
the sampler-step path.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's own CPU path on the host cores

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
# the unmodified reference travels with the repo (git-ignored, not gpurun-ignored); /root/reference is never read
_REF_LOCAL = os.path.join(ROOT, "baseline", "_ref", "reference")
if "LDX_REFERENCE_DIR" not in os.environ:
    os.environ["LDX_REFERENCE_DIR"] = _REF_LOCAL

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
REFERENCE_ARM_BUDGET_S = 150.0  # --impl reference: stop starting new images after this much timed CPU work


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


# ----------------------------------------------------------------------------------- reference legs
def reference_dir_present() -> bool:
    return os.path.exists(os.path.join(os.environ["LDX_REFERENCE_DIR"], "LightDiffusion.py"))


class ReferenceCpuLeg:
    """The UNMODIFIED reference (baseline/_ref/reference/LightDiffusion.py) on the host cores, fp32, batch 1: its
    own BaseModel / CFGGuider / ksampler("dpmpp_2m_sde", eta=0) / VAE.decode, same seeded
    weight recipe as the GPU arm."""
    kind = "reference"

    def __init__(self, threads: int):
        from lightdiffusion_b200 import sd15
        from oracle import ref_loader
        torch.set_num_threads(threads)
        R = self.R = ref_loader.load_reference(force_cpu=True)
        usd = sd15.seeded_state_dict(sd15.unet_param_shapes(), 1234)
        self.model = ref_loader.build_ref_model(R, sd15.SD15_UNET, usd, torch.float32)
        del usd
        cpu = self.cpu = torch.device("cpu")
        self.patcher = R.ModelPatcher(self.model, load_device=cpu, offload_device=cpu, current_device=cpu)
        self.vae = ref_loader.build_ref_vae(R, sd15.seeded_state_dict(sd15.vae_decoder_param_shapes(), 99), torch.float32)
        self.sigmas = R.calculate_sigmas(self.model.model_sampling, "karras", STEPS_PER_IMAGE)
        g = torch.Generator().manual_seed(7)
        self.pos = [[torch.randn(1, 77, 768, generator=g), {}]]
        self.neg = [[torch.randn(1, 77, 768, generator=g), {}]]
        self.lat = torch.zeros(1, 4, 64, 64)
        self.noise = R.prepare_noise(self.lat, 42)

    def sample(self, sigmas):
        R, p = self.R, self.patcher
        with torch.inference_mode():
            return R.sample(p, self.noise, self.pos, self.neg, CFG_SCALE, self.cpu, R.ksampler("dpmpp_2m_sde", {"eta": 0.0}),
                            sigmas, model_options=p.model_options, latent_image=self.lat, disable_pbar=True, seed=42)

    def decode(self, z):
        with torch.inference_mode():
            return self.vae.decode(z)

    def finished_image(self):
        """One whole 512x512 image: 20 guided steps + VAE decode.  Returns (seconds, detail)."""
        t0 = time.perf_counter()
        z = self.sample(self.sigmas)
        t1 = time.perf_counter()
        self.decode(z)
        t2 = time.perf_counter()
        return t2 - t0, {"t_sample_20_steps_s": round(t1 - t0, 2), "t_vae_decode_s": round(t2 - t1, 2)}

    def bounded_sample(self, n_steps: int = 3):
        """n_steps guided steps (the first sigmas of the 20-step schedule) + one full VAE decode; per-step cost is
        sigma-independent, so images/s = 1 / (20 * t_step + t_vae)."""
        t0 = time.perf_counter()
        z = self.sample(self.sigmas[:n_steps + 1])
        t1 = time.perf_counter()
        self.decode(z)
        t2 = time.perf_counter()
        t_step = (t1 - t0) / n_steps
        return 1.0 / (STEPS_PER_IMAGE * t_step + (t2 - t1)), {"t_unet_cfg_step_s": round(t_step, 3),
                                                              "t_vae_decode_s": round(t2 - t1, 3)}


class PortCpuLeg:
    """Fallback when the reference copy is not on the box: the oracle restatement of the same algorithm."""
    kind = "port"

    def __init__(self, threads: int):
        from lightdiffusion_b200 import sd15
        from oracle import ldm_oracle as O
        torch.set_num_threads(threads)
        self.O, self.sd15 = O, sd15
        self.usd = sd15.seeded_state_dict(sd15.unet_param_shapes(), 1234)
        self.vsd = sd15.seeded_state_dict(sd15.vae_decoder_param_shapes(), 99)
        g = torch.Generator().manual_seed(7)
        self.cond, self.uncond = torch.randn(1, 77, 768, generator=g), torch.randn(1, 77, 768, generator=g)
        self.noise = torch.randn(1, 4, 64, 64, generator=torch.Generator().manual_seed(42))
        from lightdiffusion_b200.pipeline import calculate_sigmas
        self.sigmas = calculate_sigmas("karras", STEPS_PER_IMAGE)

    def _run(self, sigmas):
        O, sd15 = self.O, self.sd15
        with torch.inference_mode():
            t0 = time.perf_counter()
            model = O.cfg_denoiser(self.usd, sd15.SD15_UNET, self.cond, self.uncond, CFG_SCALE)
            x = self.noise * float((1 + sigmas[0] ** 2) ** 0.5)
            z = O.sample_dpmpp_2m(model, x, sigmas) / 0.18215
            t1 = time.perf_counter()
            O.vae_decode(self.vsd, sd15.SD15_VAE, z)
            t2 = time.perf_counter()
        return t1 - t0, t2 - t1

    def finished_image(self):
        ts, tv = self._run(self.sigmas)
        return ts + tv, {"t_sample_20_steps_s": round(ts, 2), "t_vae_decode_s": round(tv, 2)}

    def bounded_sample(self, n_steps: int = 3):
        ts, tv = self._run(self.sigmas[:n_steps + 1])
        t_step = ts / n_steps
        return 1.0 / (STEPS_PER_IMAGE * t_step + tv), {"t_unet_cfg_step_s": round(t_step, 3), "t_vae_decode_s": round(tv, 3)}


def make_cpu_leg(threads: int):
    return ReferenceCpuLeg(threads) if reference_dir_present() else PortCpuLeg(threads)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores.  A step is ONE
    FINISHED 512x512 image (20 guided steps + VAE decode, batch 1 = BASELINE config 0), timed whole; ms_per_step is
    that time.  One image takes tens of seconds, so the number of timed steps is capped by a time budget and the
    line says what actually ran."""
    if env_int("RANK", 0) != 0:
        return
    cores = os.cpu_count() or 1
    leg = make_cpu_leg(cores)
    leg.sample(leg.sigmas[:2]) if hasattr(leg, "sample") else leg.bounded_sample(1)      # untimed warm-up: 1 guided step
    times, detail, t_start = [], {}, time.perf_counter()
    while len(times) < max(1, args.steps):
        if times and (time.perf_counter() - t_start) + max(times) > REFERENCE_ARM_BUDGET_S:
            break
        t, detail = leg.finished_image()
        times.append(t)
    sec = statistics.mean(times)
    value = 1.0 / sec
    sample = (f"per step: ONE finished 512x512 image, batch 1 (20 guided DPM++ 2M steps + VAE decode), fp32, "
              f"{'the unmodified reference code' if leg.kind == 'reference' else 'oracle port'}; "
              f"{len(times)} of {args.steps} requested steps fit the {REFERENCE_ARM_BUDGET_S:.0f} s budget; "
              f"warm-up = one untimed guided step")
    cfg = bench_config(args, "cpu", batch=1, gpus=1)
    cfg["workload"] = ("SD1.5 txt2img 512x512, batch 1, 20 steps DPM++ 2M Karras, CFG 7, random-init weights, fp32 on CPU "
                       "(reference path; BASELINE.json configs[0], the CPU-runnable case of the same workload)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(times), "warmup": 1, "requested_steps": args.steps, "requested_warmup": args.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": leg.kind, "sample": sample, **detail},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def gpu_reference_leg(dev, b):
    """The real bar for the kernels: the UNMODIFIED reference on the SAME GPU in the
    same process -- BaseModel fp16 through its own sample() (CFGGuider, ksampler dpmpp_2m_sde eta=0: PyTorch
    cuDNN / cuBLAS / flash-SDPA kernels) and VAE.decode in the reference's default VAE dtype -- b images, 20 steps."""
    from lightdiffusion_b200 import sd15
    from oracle import ref_loader
    R = ref_loader.load_reference()
    usd = sd15.seeded_state_dict(sd15.unet_param_shapes(), 1234, dtype=torch.float16, device=dev)
    m16 = ref_loader.build_ref_model(R, sd15.SD15_UNET, usd, torch.float16).to(dev)
    del usd
    patcher = R.ModelPatcher(m16, load_device=dev, offload_device=dev, current_device=dev)
    vae = ref_loader.build_ref_vae(R, sd15.seeded_state_dict(sd15.vae_decoder_param_shapes(), 99), R.VAE_DTYPE, device=dev)
    sigmas = R.calculate_sigmas(m16.model_sampling, "karras", STEPS_PER_IMAGE)
    g = torch.Generator().manual_seed(7)
    pos, neg = [[torch.randn(b, 77, 768, generator=g), {}]], [[torch.randn(b, 77, 768, generator=g), {}]]
    lat = torch.zeros(b, 4, 64, 64)
    noise = R.prepare_noise(lat, 42)

    def once():
        with torch.inference_mode():
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            z = R.sample(patcher, noise, pos, neg, CFG_SCALE, dev, R.ksampler("dpmpp_2m_sde", {"eta": 0.0}), sigmas,
                         model_options=patcher.model_options, latent_image=lat, disable_pbar=True, seed=42)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            img = vae.decode(z)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
        assert tuple(img.shape) == (b, 512, 512, 3)
        return t1 - t0, t2 - t1

    once()
    runs = [once() for _ in range(3)]
    ts, tv = statistics.median(r[0] for r in runs), statistics.median(r[1] for r in runs)
    del m16, vae, patcher
    torch.cuda.empty_cache()
    return {"impl": "unmodified reference, PyTorch " + torch.__version__ + " on the same GPU, same run",
            "dtype": f"fp16 UNet, {str(R.VAE_DTYPE).split('.')[-1]} VAE (the reference's defaults)",
            "images_per_gpu": b, "value": b / (ts + tv), "unit": UNIT, "sample_s": round(ts, 4),
            "unet_step_ms": round(ts / STEPS_PER_IMAGE * 1e3, 3), "vae_decode_s": round(tv, 4),
            "timing": "wall clock around sample() and VAE.decode() with torch.cuda.synchronize() on both sides, "
                      "median of 3 after 1 warm-up; VAE.decode returns host memory (D2H inside, like our e2e)"}


def bench_config(args, where, batch=None, gpus=None):
    batch = args.batch if batch is None else batch
    gpus = args.gpus if gpus is None else gpus
    return {"workload": f"SD1.5 txt2img 512x512, batch {batch} per GPU, fp16, 20 steps DPM++ 2M Karras, CFG 7"
                        + (", images sharded over GPUs" if gpus > 1 else ", on 1 B200"),
            "images_per_gpu": batch, "global_batch": batch * gpus, "latent": "64x64",
            "sampler_steps": STEPS_PER_IMAGE, "cfg_scale": CFG_SCALE, "scheduler": "karras",
            "weights": "random-init SD1.5 architecture (seeded)", "parallelism": f"images sharded dp{gpus}",
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
    scaling = "weak"
    if args.total_images:                       # BASELINE config 3 as written: a fixed total sharded over the ranks
        assert args.total_images % world == 0, "--total-images must be a multiple of the GPU count"
        args.batch = args.total_images // world
        scaling = "strong"
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
    # N = 1: the fp32 [b, 512, 512, 3] images VAE.decode returns; N > 1: uint8 images (what the reference saves),
    # gathered to rank 0 only, which reads the whole job's stack back
    multi = world > 1
    h_img = (torch.empty(b * world if rank == 0 else 0, 512, 512, 3, dtype=torch.uint8) if multi
             else torch.empty(b, 512, 512, 3, dtype=torch.float32)).pin_memory()

    def step_resident():
        z, img = pipe.generate(d_noise, d_cond, d_uncond, CFG_SCALE, sig, "dpmpp_2m", u8=multi)
        return gather_images(img, b * world) if multi else img

    def step_e2e():
        n, c, u = h_noise.to(dev, non_blocking=True), h_cond.to(dev, non_blocking=True), h_uncond.to(dev, non_blocking=True)
        z, img = pipe.generate(n, c, u, CFG_SCALE, sig, "dpmpp_2m", u8=multi)
        if multi:
            img = gather_images(img, b * world)
        if img is not None:
            h_img.copy_(img, non_blocking=True)

    def timed(fn, k):
        if multi:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if multi:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if multi:
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

    # UNet step time (graph replay of the 2b-sample CFG evaluation + its tail) and per-kernel roofline
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
    t_kind, n_kind, f_kind = {}, {}, {}
    for kind, flops, tag, t in per_op:
        t_kind[kind] = t_kind.get(kind, 0.0) + t
        n_kind[kind] = n_kind.get(kind, 0) + 1
        f_kind[kind] = f_kind.get(kind, 0.0) + flops        # FLOPs the launches execute (Program.meta)
    t_total = sum(t_kind.values())
    # the tensor-core GEMM family: ldx_gemm_kernel (linears, strided / up / 8x8 convs) and ldx_conv_halo_kernel (the
    # ResBlock 3x3 convs with GroupNorm + SiLU in the operand path); the roofline line is for the family, the
    # per-kernel split follows in roofline["kernels"]
    # ... and ldx_ff_fused_kernel (GEGLU projection + FF2 of the 320-channel blocks in one launch): its time and FLOPs
    # belong to the same family, the algorithmic figure counts those two GEMMs
    fam = ("gemm", "conv_halo", "ff_fused")
    kname = {"gemm": "ldx_gemm_kernel", "conv_halo": "ldx_conv_halo_kernel", "ff_fused": "ldx_ff_fused_kernel"}
    gemm_ms = sum(t_kind.get(k, 0.0) for k in fam)
    gemm_flops = sum(f_kind.get(k, 0.0) for k in fam)
    gemm_n = sum(n_kind.get(k, 0) for k in fam)
    # algorithmic = the reference's work for 2b samples through every layer (contractual figure); executed = what
    # the launches issue: the up-conv fold, the hoisted context K/V and the shared CFG prefix remove work, the
    # fused-residual identity K steps and 64-channel K padding add some
    achieved = GF_UNET_GEMM_SAMPLE * 1e9 * 2 * b / (gemm_ms * 1e-3) / 1e12
    executed = gemm_flops / (gemm_ms * 1e-3) / 1e12
    traffic, traffic_src = None, None
    for name in ("r02c_gemm_traffic.json", "r02b_gemm_traffic.json", "r02_gemm_traffic.json", "r01_gemm_traffic.json"):   # committed ncu measurement (per launch)
        tpath = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), f"profiles/{name}: " + tj.get("source", "")
            break
    roofline = {"kernel": "ldx_gemm_kernel + ldx_conv_halo_kernel + ldx_ff_fused_kernel (implicit-GEMM linear / conv / fused "
                          "feed-forward, tcgen05)", "bound": "tensor",
                "achieved": achieved, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["tflops_sustained"], "traffic": traffic, "traffic_source": traffic_src,
                "achieved_executed": executed, "frac_executed": executed / peaks["tflops_sustained"],
                "executed_gflop_per_unet_step": gemm_flops / 1e9,
                "kernels": {kname[k]:
                            {"ms": round(t_kind.get(k, 0.0), 4), "launches": n_kind.get(k, 0),
                             "executed_tflops": round(f_kind.get(k, 0.0) / max(t_kind.get(k, 0.0), 1e-9) / 1e9, 1)}
                            for k in fam if k in t_kind},
                "algorithmic_gflop_per_unet_step": GF_UNET_GEMM_SAMPLE * 2 * b,
                "peak_source": peaks["source"] + ", sustained bf16 cuBLAS figure (kernel timed inside a long step)",
                "launches_per_unet_step": gemm_n, "ms_per_unet_step": gemm_ms,
                "share_of_unet_step": gemm_ms / t_total,
                "algorithmic_gflop_per_launch_avg": GF_UNET_GEMM_SAMPLE * 2 * b / max(1, gemm_n),
                "frac_of_nominal_2250": achieved / 2250.0}
    breakdown = {k: {"ms": round(v, 4), "launches": n_kind[k], "share": round(v / t_total, 4)} for k, v in t_kind.items()}
    if "attention" in t_kind:
        a_ms = t_kind["attention"] * 1e-3
        breakdown["attention"]["tflops"] = round(GF_UNET_ATTN_SAMPLE * 1e9 * 2 * b / a_ms / 1e12, 1)
        breakdown["attention"]["tflops_executed"] = round(f_kind.get("attention", 0.0) / a_ms / 1e12, 1)

    cpu_baseline = None
    gpu_reference = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        leg = make_cpu_leg(cores)
        v, detail = leg.bounded_sample(3)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": leg.kind,
                        "sample": "3 guided DPM++ 2M steps (CFG UNet evaluations, batch 1, 64x64 latent, fp32) through "
                                  + ("the unmodified reference's sample()" if leg.kind == "reference" else "the oracle port")
                                  + " + one full VAE decode; images/s = 1/(20*t_step + t_vae)", **detail}
        del leg
    if rank == 0 and world == 1 and not args.no_gpu_reference and reference_dir_present():
        try:
            gpu_reference = gpu_reference_leg(dev, b)
            gpu_reference["speedup_e2e"] = round(e2e_value / gpu_reference["value"], 3)
            gpu_reference["speedup_unet_step"] = round(gpu_reference["unet_step_ms"] / unet_ms, 3)
        except Exception as exc:                      # the reference arm must never take the product line down
            gpu_reference = {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}

    if rank == 0:
        cfg = bench_config(args, "gpu")
        if args.total_images:
            cfg["workload"] = (f"SD1.5 txt2img 512x512, batch {args.total_images} total, images sharded across "
                               f"{world} B200 ({b} per GPU), fp16, 20 steps DPM++ 2M Karras, CFG 7")
        if multi:
            cfg["gather"] = "uint8 [b,512,512,3] per rank -> rank 0 (NCCL gather), inside the timed region"
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": scaling, "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
                "config": cfg,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches_per_step * args.steps,
                "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline, "gpu_reference": gpu_reference,
                "unet_step_ms": unet_ms, "unet_step_tflops": GF_UNET_SAMPLE * 1e9 * 2 * b / (unet_ms * 1e-3) / 1e12,
                "unet_step_tflops_executed": (gemm_flops + f_kind.get("attention", 0.0)) / (unet_ms * 1e-3) / 1e12,
                "pipeline_tflops_per_gpu": TF_IMAGE * value / world,
                "frac_of_sustained_peak": TF_IMAGE * value / world / peaks["tflops_sustained"],
                "unet_breakdown": breakdown}
        print(json.dumps(line), flush=True)
    if multi:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="images per GPU (BASELINE.json config 2: 8)")
    ap.add_argument("--total-images", type=int, default=0,
                    help="fixed total sharded over the ranks (BASELINE.json config 3: 64); overrides --batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
