"""Host-side logic that needs no GPU: architecture schema, schedules, sampler coefficients, sharding and
the N>1 gather over gloo (world_size 2)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from lightdiffusion_b200 import sd15
from lightdiffusion_b200.pipeline import calculate_sigmas, shard_range, sigmas_for_denoise
from lightdiffusion_b200.samplers import dpmpp2m_coefficients, get_ancestral_step


def test_schema_counts():
    assert sd15.count_params(sd15.unet_param_shapes()) == 859520964
    lay = sd15.unet_layout(sd15.SD15_UNET)
    assert len(lay.input_blocks) == 12 and len(lay.output_blocks) == 12 and len(lay.middle) == 3
    n_res = sum(isinstance(s, sd15.ResSpec) for b in lay.input_blocks + [lay.middle] + lay.output_blocks for s in b)
    n_att = sum(isinstance(s, sd15.AttnSpec) for b in lay.input_blocks + [lay.middle] + lay.output_blocks for s in b)
    assert (n_res, n_att) == (22, 16)                      # SURVEY.md App. A module census
    dec = sd15.vae_decoder_param_shapes()
    assert sum(torch.Size(s).numel() for k, s in dec.items() if k.startswith("decoder.")) == 49490179


def test_seeded_weights_are_reproducible():
    shapes = sd15.unet_param_shapes(sd15.tiny_unet_config())
    a, b = sd15.seeded_state_dict(shapes, 5), sd15.seeded_state_dict(shapes, 5)
    assert all(torch.equal(a[k], b[k]) for k in a)
    c = sd15.seeded_state_dict(shapes, 6)
    assert not torch.equal(a["time_embed.0.weight"], c["time_embed.0.weight"])


def test_schedules_match_reference_golden(golden):
    g = golden("schedules.pt")
    assert torch.equal(calculate_sigmas("karras", 20), g["karras20"])
    assert torch.equal(calculate_sigmas("karras", 8), g["karras8"])
    assert torch.equal(calculate_sigmas("normal", 10), g["normal10"])
    assert torch.equal(calculate_sigmas("normal", 22), g["normal22"])
    assert torch.equal(sigmas_for_denoise("normal", 4, 0.45), golden("samplers_tiny.pt")["sigmas2"])


def test_dpmpp2m_coefficients_reproduce_reference_trajectory(golden):
    """x <- c_x x + c_d D + c_o (D - D_old) with our host coefficients == the real reference's
    sample_dpmpp_2m_sde(eta=0) trajectory (golden) on the closed-form denoiser."""
    g = golden("samplers_toy.pt")
    sig, x = g["sigmas"], g["x0"].clone()
    old, h_last = None, None
    for i in range(len(sig) - 1):
        d = x / (1.0 + sig[i] ** 2)
        if sig[i + 1] == 0:
            x, h = d, None
        else:
            c_x, c_d, c_o, h = dpmpp2m_coefficients(sig, i, h_last if old is not None else None)
            x = c_x * x + c_d * d + (0 if c_o is None else c_o * (d - old))
        old, h_last = d, h
    assert float((x - g["dpmpp2m"]).norm() / g["dpmpp2m"].norm()) < 2e-6
    sd, su = get_ancestral_step(sig[3], sig[4])
    assert abs(float(sd ** 2 + su ** 2) - float(sig[4] ** 2)) < 1e-5


def test_neighbour_component_schemas(golden):
    """Parameter schemas of the components around the loop, counted against the reference modules (fixtures carry
    the reference's own parameter counts) and against each other."""
    enc = sd15.vae_encoder_param_shapes()
    assert sum(torch.Size(s).numel() for k, s in enc.items() if k.startswith("encoder.")) == golden("vae_encode.pt")["n_encoder_params"]
    clip = sd15.clip_text_param_shapes()
    assert sd15.count_params(clip) == 49408 * 768 + 77 * 768 + 12 * (4 * (768 * 768 + 768) + 2 * 2 * 768
                                                                   + 3072 * 768 + 3072 + 768 * 3072 + 768) + 2 * 768 + 768 * 768
    from lightdiffusion_b200.esrgan import esrgan_param_shapes, to_old_arch
    for scale in (2, 4):
        old = esrgan_param_shapes(3, scale, arch="old")
        for arch in ("esrgan", "realesrgan"):
            new = esrgan_param_shapes(3, scale, arch=arch)
            mapped = to_old_arch({k: torch.empty(0).new_empty(v) for k, v in new.items()})
            assert set(mapped) == set(old) and all(tuple(mapped[k].shape) == old[k] for k in old), (scale, arch)
    with pytest.raises(Exception):
        to_old_arch({"foo.weight": torch.zeros(1)})
    from lightdiffusion_b200.taesd import taesd_param_shapes
    assert sd15.count_params(taesd_param_shapes("decoder")) == 4 * 64 * 9 + 64 + 10 * 3 * (64 * 64 * 9 + 64) + 3 * 64 * 64 * 9 + 64 * 3 * 9 + 3
    assert sd15.count_params(taesd_param_shapes("encoder")) == 3 * 64 * 9 + 64 + 10 * 3 * (64 * 64 * 9 + 64) + 3 * 64 * 64 * 9 + 64 * 4 * 9 + 4


def test_host_side_tables_and_controllers():
    """bislerp axis tables, the empty-prompt token row and the PID step-size controller against the restatements."""
    from oracle import ldm_oracle as O
    from lightdiffusion_b200.upscale import _axis_table_cpu
    for old, new in ((64, 128), (12, 18), (12, 17), (8, 4), (5, 5), (1, 3)):
        r, i1, i2 = _axis_table_cpu(old, new)
        ro, o1, o2 = O._bilinear_table(old, new)
        assert torch.equal(r, ro) and torch.equal(i1.long(), o1) and torch.equal(i2.long(), o2)
        assert int(i2.max()) <= old - 1 and int(i1.min()) >= 0
    from lightdiffusion_b200.clip import SD1ClipEncoder, gen_empty_tokens
    assert gen_empty_tokens(SD1ClipEncoder.special_tokens, 77) == [49406, 49407] + [49407] * 75
    from lightdiffusion_b200.samplers import PIDStepSizeController
    pid = PIDStepSizeController(0.05, 0.1, 1.0, 0.05, order=3, accept_safety=0.81)
    accepts = [pid.propose_step(e) for e in (0.5, 0.9, 3.0, 40.0, 0.2, 1.0)]
    assert accepts == [True, True, False, False, True, True]      # large errors are rejected and shrink h
    assert 0.0 < pid.h < 0.2
    pid2 = PIDStepSizeController(0.05, 0.0, 1.0, 0.0, order=3)
    pid2.propose_step(1.0)
    assert abs(pid2.h - 0.05) < 1e-6                              # unit error leaves the step size unchanged


def test_pid_controller_matches_three_slot_restatement_bit_for_bit():
    """The product's two-entry-history controller against the reference's three-slot bookkeeping (LD.py:944-973, as restated
    inline in oracle/ldm_oracle.py's dpm_adaptive): identical accept sequences and bit-identical step sizes on random error
    sequences, including rejected first steps and long reject runs."""
    import math
    import random
    from lightdiffusion_b200.samplers import PIDStepSizeController

    def three_slot(h, p, i, d, order, safety, errors, eps=1e-8):
        b1, b2, b3 = (p + i + d) / order, -(p + 2 * d) / order, d / order
        errs, out = [], []
        for e in errors:
            inv = 1 / (float(e) + eps)
            if not errs:
                errs = [inv, inv, inv]
            errs[0] = inv
            f = 1 + math.atan(errs[0] ** b1 * errs[1] ** b2 * errs[2] ** b3 - 1)
            ok = f >= safety
            if ok:
                errs[2], errs[1] = errs[1], errs[0]
            h *= f
            out.append((ok, h))
        return out

    rng = random.Random(7)
    for trial in range(200):
        p, i, d = rng.uniform(0, 0.3), rng.uniform(0.3, 1.2), rng.uniform(0, 0.2)
        order, safety = rng.choice([1, 1.5, 2, 3]), rng.choice([0.81, 0.9, 0.5])
        errors = [10 ** rng.uniform(-3, 2.5) for _ in range(rng.randint(1, 40))]
        if trial % 5 == 0:
            errors[0] = 500.0                   # rejected first proposal: the history stays seeded with it
        pid = PIDStepSizeController(0.05, p, i, d, order, safety)
        got = [(pid.propose_step(e), pid.h) for e in errors]
        assert got == three_slot(0.05, p, i, d, order, safety, errors)


def test_shard_range_partitions_everything():
    for n, w in [(64, 8), (64, 2), (10, 4), (3, 8), (1, 1)]:
        got = []
        for r in range(w):
            lo, hi = shard_range(n, r, w)
            got += list(range(lo, hi))
        assert got == list(range(n))


def _gather_worker(rank, world, port, n_total, q):
    import torch.distributed as dist
    from lightdiffusion_b200.pipeline import gather_images, shard_range
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_total, rank, world)
    local = torch.stack([torch.full((4, 4, 3), float(i)) for i in range(lo, hi)]) if hi > lo else torch.zeros(0, 4, 4, 3)
    out = gather_images(local, n_total)
    u8 = gather_images((local % 251).to(torch.uint8), n_total)         # the compact form the bench gathers
    if rank == 0:
        assert u8.dtype == torch.uint8 and u8[:, 0, 0, 0].tolist() == [i % 251 for i in range(n_total)]
        q.put(out[:, 0, 0, 0].tolist())
    else:
        assert out is None and u8 is None                              # the images land on rank 0 only
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [6, 5])
def test_gather_images_world2_gloo(n_total):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == [float(i) for i in range(n_total)]
