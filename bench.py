#!/usr/bin/env python
"""Benchmark of the ControlNet-SD1.5 denoising hot path (BASELINE.json configs[1]):
256x384 image (latent 32x48), batch 1, DDIM 20 steps, CFG 9.0, bf16, synthetic data / random-init weights.

A "step" is ONE DDIM denoising step = 2 x (ControlNet + UNet) forward (cond + uncond, run as one batch of 2) + CFG
combine + DDIM x_{t-1} update. Steps run in 20-step images; between images the latent is rewound on the device.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl sdeo|reference]

value  : device-resident throughput, steps/s over all ranks (inputs already in HBM, CUDA-event timed, max over ranks)
e2e    : same metric through the public API — DDIMSampler.sample(...) fed pinned HOST tensors (x_T, hint, contexts
         copied host->device every image, final latents copied device->host), ceil(K/20) images
roofline: the dominant kernel (conv_gemm_kernel, tensor-core bound) — executed conv/linear FLOPs of one step divided
         by the time those launches take INSIDE the timed step graph: the library's kernel trace (globaltimer stamps
         written by block 0 of every kernel during one graph replay: dependency resolved -> block end, plus the idle
         tail that follows each conv kernel before any other kernel is past its dependency). The older figure —
         CUDA events around every eager launch on a back-logged stream — is kept as `achieved_eager_events`: it
         serialises the step and adds ~10 us of event/launch latency to every 5-10 us kernel.
cpu_baseline / --impl reference: the CPU fp32 port of the reference path (oracle/) timed on the host cores.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

S_DDIM = 20
LATENT_HW = (32, 48)
CFG_SCALE = 9.0
WORKLOAD = "ControlNet-canny SD1.5 256x384 batch 1, DDIM 20 steps, CFG 9.0 (BASELINE configs[1])"
FLOPS_PER_STEP = 740.0e9          # one DDIM step = 2 x (ControlNet 95.6 + UNet 274.4) GF at 32x48
WEIGHT_BYTES_PER_STEP = 2.442e9   # bf16 UNet 1.719 GB + ControlNet 0.723 GB, streamed once per step (cond+uncond batched)
# BASELINE.md §2 (forward hooks on the reference's modules): FLOPs of ONE DDIM step = 2 x (ControlNet + UNet) for ONE
# image, and of one VAE decode, per latent size
STEP_FLOPS = {(32, 48): 740.0e9, (64, 64): 2173.2e9, (96, 96): 5843.2e9}
VAE_FLOPS = {(32, 48): 934.9e9, (64, 64): 2514.5e9, (96, 96): 5754.3e9}
VAE_GN_ELEMS = {(32, 48): 172.2e6, (64, 64): 459.3e6, (96, 96): 1033.4e6}   # GroupNorm elements per decode (4 B each in bf16)


def conv_dram_bytes_per_launch():
    """dram__bytes_read.sum + dram__bytes_write.sum per conv_gemm_kernel launch, from the newest committed ncu launch list
    of the step graph (profiles/*launches_step_summary.txt, written by tools/ncu_launch_summary.py from the ncu CSV of
    `tools/profile_step.py --graph`). None if no such profile is committed."""
    import glob
    import re
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*launches_step_summary.txt"))):
        with open(path) as f:
            m = re.search(r"conv_gemm_kernel DRAM traffic per launch: (\d+) bytes", f.read())
        if m:
            best = (int(m.group(1)), os.path.relpath(path, ROOT))
    return best


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(bf16_burst=p.get("bf16_tflops", 1590.0), bf16_sustained=p.get("bf16_tflops_sustained", 1400.0),
                    hbm_gbs=p.get("hbm_gbs", 6650.0), source="measured (MEASURED_PEAKS.json)")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


def host_inputs(pin, latent_hw=LATENT_HW, batch=1):
    """Synthetic inputs: x_T seed 2946901; contexts seeds 1/2; a binary edge map as the hint."""
    h, w = latent_hw
    x_T = torch.randn((batch, 4, h, w), generator=torch.Generator().manual_seed(2946901))
    ctx_c = torch.randn((batch, 77, 768), generator=torch.Generator().manual_seed(1))
    ctx_u = torch.randn((batch, 77, 768), generator=torch.Generator().manual_seed(2))
    r = torch.rand((batch, 1, 8 * h, 8 * w), generator=torch.Generator().manual_seed(7))
    hint = (r > 0.9).float().expand(-1, 3, -1, -1).contiguous()
    ts = [x_T, ctx_c, ctx_u, hint]
    if pin:
        ts = [t.pin_memory() for t in ts]
    return ts


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().strip().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# CPU reference arm (the oracle port of the reference's PyTorch path)
# ---------------------------------------------------------------------------------------------------------------
def cpu_reference_steps(max_steps, budget_s):
    """Runs DDIM steps of the CPU fp32 port on all host cores. Returns (steps_run, seconds, cores)."""
    from oracle import sd15_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.SD15
    sd_unet = O.make_weights(O.unet_param_spec(cfg), seed=1234, prefix="unet.")
    sd_cn = O.make_weights(O.controlnet_param_spec(cfg), seed=1234, prefix="control.")
    x, cond, uncond = O.make_inputs(cfg, 1, *LATENT_HW)
    sch = O.ddim_schedule(S_DDIM)
    steps = list(reversed(sch["timesteps"].tolist()))

    def one_step(x, i):
        index = S_DDIM - (i % S_DDIM) - 1
        ts = torch.full((1,), int(steps[i % S_DDIM]), dtype=torch.long)
        e_c = O.apply_model(sd_unet, sd_cn, cfg, x, ts, cond)
        e_u = O.apply_model(sd_unet, sd_cn, cfg, x, ts, uncond)
        e = e_u + CFG_SCALE * (e_c - e_u)
        return O.ddim_update(x, e, float(sch["alphas"][index]), float(sch["alphas_prev"][index]), 0.0,
                             float(sch["sqrt_one_minus_alphas"][index]))[0]

    with torch.no_grad():
        t0 = time.time()
        x = one_step(x, 0)  # warm-up (thread pools, allocator)
        t_warm = time.time() - t0
        n = max(1, min(max_steps, int(budget_s / max(t_warm, 1e-3))))
        t0 = time.time()
        for i in range(n):
            x = one_step(x, i + 1)
        dt = time.time() - t0
    assert torch.isfinite(x).all()
    return n, dt, cores


def main_config(world):
    """The `config` object of BOTH arms (same keys and values, so the driver's same_config check compares like with like)."""
    return {"workload": WORKLOAD, "latent": list(LATENT_HW), "ddim_steps": S_DDIM, "cfg_scale": CFG_SCALE,
            "batch_per_gpu": 1, "parallelism": f"replicas x{world} (one image stream per GPU, no collective in the loop)",
            "l2_policy": "no flush needed: 2.44 GB of bf16 weights are streamed every step (>> 126 MB L2)",
            "weights": "random-init (N(0, var) per SURVEY 8d)"}


def run_reference_arm(args, rank):
    if rank != 0:
        return
    n, dt, cores = cpu_reference_steps(args.steps, budget_s=150.0)
    v = n / dt
    line = {
        "impl": "reference", "metric": "denoise_steps_per_s", "value": v, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / v, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": main_config(args.gpus),
        "cpu_baseline": {"value": v, "unit": "steps/s", "cores": cores, "kind": "port",
                         "sample": f"{n} of {args.steps} DDIM steps executed on the host CPU (each = 2 x (ControlNet+UNet) fp32 "
                                   f"+ CFG + DDIM update, oracle port of the reference modules; 1 extra warm-up step)"},
        "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def conv_roofline(eng, pk):
    """Eager pass of one step with CUDA events around every conv_gemm launch, on a back-logged stream (a spin kernel
    keeps the GPU busy while the host enqueues, so event deltas are kernel durations, not launch gaps)."""
    from stablediffusioneo_b200 import ops
    rec = []
    orig = ops.conv2d

    def timed(x, pw, *a, **kw):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        y = orig(x, pw, *a, **kw)
        e.record()
        n, h, w, _ = x.shape
        stride = kw.get("stride", 1)
        k = pw.ksize
        ho, wo = (h + 2 * (k // 2) - k) // stride + 1, (w + 2 * (k // 2) - k) // stride + 1
        flops = 2.0 * n * ho * wo * pw.cout * (pw.c1 + pw.c2) * k * k
        rec.append((s, e, flops))
        return y

    eng.reset_latent()
    torch.cuda.synchronize()
    ops.conv2d = timed
    try:
        torch.cuda._sleep(int(1.5e9))  # ~1 s of GPU spin: the host runs ahead and queues the whole step behind it
        eng._step()
        torch.cuda.synchronize()
    finally:
        ops.conv2d = orig
    eng.reset_latent()
    total_ms = sum(s.elapsed_time(e) for s, e, _ in rec)
    total_flops = sum(f for _, _, f in rec)
    eager = total_flops / (total_ms * 1e-3) / 1e12
    # in-graph: one replay of the captured step with the kernel trace on
    from stablediffusioneo_b200 import trace
    for _ in range(3):
        eng.step()
    recs = trace.capture(eng.step, eng.x_lat.device)
    eng.reset_latent()
    summ = trace.summarize(recs)
    conv = summ["kinds"]["conv"]
    conv_us = conv["busy_us"] + conv["tail_us"]
    busy_all = sum(k["busy_us"] for k in summ["kinds"].values())
    achieved = total_flops / (conv_us * 1e-6) / 1e12
    return {"bound": "tensor", "kernel": "conv_gemm_kernel", "achieved": achieved, "peak": pk["bf16_sustained"],
            "unit": "TFLOP/s", "frac": achieved / pk["bf16_sustained"],
            # dram__bytes_read+write per launch, averaged over the conv launches of one step (ncu --graph-profiling node,
            # profiles/r01c_launches_step_summary.txt; cold-cache replays; algorithmic: 2.442 GB of weights / launches)
            "traffic": (conv_dram_bytes_per_launch() or (None, None))[0],
            "traffic_unit": "bytes/launch (ncu dram__bytes_read+write, cold cache)",
            "traffic_source": (conv_dram_bytes_per_launch() or (None, None))[1],
            "launches_per_step": conv["launches"], "flops_per_step": total_flops,
            "kernel_ms_per_step": conv_us * 1e-3, "kernel_busy_ms_per_step": conv["busy_us"] * 1e-3,
            "avg_launch_us": conv_us / conv["launches"],
            # the same FLOPs over the wall-clock time during which at least one conv kernel runs (+ tails): the two
            # streams' launches overlap in time, `achieved` charges every launch its full duration
            "achieved_while_running": total_flops / ((conv["union_us"] + conv["tail_us"]) * 1e-6) / 1e12,
            "kernel_share_of_step_kernel_time": conv["busy_us"] / busy_all,
            "method": "sdeo_set_trace during one replay of the timed step graph: sum over conv launches of (grid "
                      "dependency resolved -> block-0 end) + the idle tail after each until another kernel runs",
            "achieved_eager_events": eager, "eager_events_ms_per_step": total_ms,
            "step_trace": {k: {"launches": v["launches"], "busy_us": round(v["busy_us"], 1), "tail_us": round(v["tail_us"], 1)}
                           for k, v in summ["kinds"].items()},
            "step_trace_span_us": summ["span_us"], "step_trace_busy_union_us": summ["busy_union_us"],
            "peak_source": pk["source"] + ", sustained figure (kernel timed inside a long step)"}


def timed_steps(eng, steps, warmup, dist, dev, min_seconds=1.0, clocks_for=None):
    """K device-resident steps of `eng` (20-step images, latent rewound on the device between images), repeated R times
    so that the timed region lasts >= min_seconds (the same K-step block every repeat). CUDA events on the launching
    stream, barrier + synchronize on both sides, max over ranks. Returns (ms per K-step block, R, eager launches, clocks)."""
    from stablediffusioneo_b200 import ops
    pos = [0]

    def run_steps(k):
        for _ in range(k):
            if pos[0] % S_DDIM == 0:
                eng.reset_latent()
            eng.step()
            pos[0] += 1

    eng.reset_latent()
    run_steps(max(warmup, 3))
    pos[0] = 0
    # size the repeat count from a short probe
    s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s0.record()
    run_steps(3)
    e0.record()
    torch.cuda.synchronize()
    est_ms = max(s0.elapsed_time(e0) / 3.0, 1e-3)
    repeats = max(1, int(math.ceil(min_seconds * 1e3 / (est_ms * steps))))
    if dist is not None:
        t = torch.tensor([repeats], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        repeats = int(t.item())
    pos[0] = 0
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    clocks = ClockSampler(clocks_for) if clocks_for is not None else None
    n0 = ops.LAUNCHES
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    start.record()
    run_steps(steps * repeats)
    end.record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    elapsed_ms = start.elapsed_time(end)
    eager = ops.LAUNCHES - n0
    if dist is not None:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    return elapsed_ms / repeats, repeats, eager, (clocks.stop() if clocks is not None else None)


def denoise_config(model, latent_hw, batch, steps, warmup, dist, dev, world, pk, name, guess=False):
    """One of the other BASELINE configs as a device-resident + end-to-end measurement: builds its own engine.
    guess: hackathon.process(guess_mode=True) (canny2image_torch.py:48,54): no hint on the unconditional branch (the ControlNet
    runs on the conditional rows only), control strengths 0.825^(12-i)."""
    from stablediffusioneo_b200.cldm.ddim_hacked import DDIMSampler
    sampler = DDIMSampler(model)
    x_T, ctx_c, ctx_u, hint = host_inputs(pin=True, latent_hw=latent_hw, batch=batch)
    cond = {"c_concat": [hint], "c_crossattn": [ctx_c]}
    uncond = {"c_concat": None if guess else [hint], "c_crossattn": [ctx_u]}
    saved_scales = list(model.control_scales)
    if guess:
        model.control_scales = [1.0 * (0.825 ** float(12 - i)) for i in range(13)]

    def sample():
        out, _ = sampler.sample(S_DDIM, batch, (4,) + tuple(latent_hw), cond, verbose=False, eta=0.0, x_T=x_T,
                                unconditional_guidance_scale=CFG_SCALE, unconditional_conditioning=uncond)
        return out.to("cpu")

    out = sample()
    assert torch.isfinite(out).all()
    sample()
    eng = sampler._engine
    ms_block, repeats, _, _ = timed_steps(eng, steps, warmup, dist, dev, min_seconds=0.5)
    ms_step = ms_block / steps
    n_img = 2
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n_img):
        sample()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    flops = STEP_FLOPS[tuple(latent_hw)] * batch
    tf = flops / (ms_step * 1e-3) / 1e12
    res = {"name": name, "latent": list(latent_hw), "batch_per_gpu": batch, "n_gpus": world,
           "steps_per_s": world * 1e3 / ms_step, "ms_per_step": ms_step,
           "images_per_s": world * batch * 1e3 / (ms_step * S_DDIM), "timed_steps": steps * repeats,
           "launches_per_step": getattr(eng, "launches_per_step", None),
           "e2e": {"steps_per_s": world * n_img * S_DDIM / e2e_s, "images_per_s": world * n_img * batch / e2e_s, "images": n_img * batch,
                   "h2d_bytes_per_step": (x_T.numel() + ctx_c.numel() + ctx_u.numel() + 2 * hint.numel()) * 4 / S_DDIM,
                   "d2h_bytes_per_step": x_T.numel() * 4 / S_DDIM},
           "roofline": {"bound": "tensor", "achieved": tf, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                        "frac": tf / pk["bf16_sustained"], "flops_per_step": flops,
                        "note": "whole step (all kernels) vs BASELINE.md section 2 FLOPs; sustained cuBLAS bf16 peak"}}
    del sampler._engine
    sampler._engine = None
    model.control_scales = saved_scales
    torch.cuda.empty_cache()
    return res


def vae_config(model, latent_hw, batch, dev, pk, name, iters=5):
    """VAE decode (latents resident -> uint8 NHWC image on the device) and end to end (host latents -> host image)."""
    h, w = latent_hw
    z = (torch.randn((batch, 4, h, w), generator=torch.Generator().manual_seed(11)) * 0.18215 * 4.0).pin_memory()
    zd = z.to(dev)
    for _ in range(4):   # (calls 1-2 run eagerly and tune the conv shapes, call 3 captures the decode graph)
        u8 = model.decode_first_stage_u8(zd)
    torch.cuda.synchronize()
    assert u8.shape == (batch, 8 * h, 8 * w, 3)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        u8 = model.decode_first_stage_u8(zd)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    t0 = time.perf_counter()
    for _ in range(2):
        img = model.decode_first_stage_u8(z.to(dev, non_blocking=True)).cpu()
    e2e_s = (time.perf_counter() - t0) / 2
    flops = VAE_FLOPS[tuple(latent_hw)] * batch
    gn_bytes = VAE_GN_ELEMS[tuple(latent_hw)] * batch * 4.0
    tf = flops / (ms * 1e-3) / 1e12
    return {"name": name, "latent": list(latent_hw), "batch": batch, "images_per_s": batch * 1e3 / ms, "ms_per_batch": ms,
            "e2e": {"images_per_s": batch / e2e_s, "h2d_bytes": z.numel() * 4, "d2h_bytes": img.numel()},
            "roofline": {"bound": "tensor", "achieved": tf, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": tf / pk["bf16_sustained"], "flops_per_batch": flops,
                         "groupnorm_bytes": gn_bytes, "groupnorm_hbm_floor_ms": gn_bytes / (pk["hbm_gbs"] * 1e9) * 1e3,
                         "note": "whole decode vs BASELINE.md section 2 FLOPs; the GroupNorm floor is 4 B/element at the "
                                 "measured HBM copy bandwidth"}}


def other_configs(model, args, dist, dev, world, pk, rank):
    """The other BASELINE.json configs (and batch > 1 images/s at 256x384) measured in the same run. N > 1: only
    configs[2] (512x512, one image per GPU, sharded over the ranks). Each entry is guarded: a failure is reported in
    place and never costs the headline line."""
    jobs = [("configs[2] ControlNet-canny SD1.5 512x512, 1 image per GPU, DDIM 20, CFG 9", "denoise", (64, 64), 1, 20)]
    if world == 1:
        jobs += [("configs[3] SD1.5 UNet+ControlNet 768x768 batch 4, DDIM 20, CFG 9", "denoise", (96, 96), 4, 10),
                 ("256x384 batch 4 per GPU (images/s)", "denoise", (32, 48), 4, 20),
                 ("256x384 batch 8 per GPU (images/s)", "denoise", (32, 48), 8, 20),
                 ("256x384 batch 1, guess mode (no ControlNet on the unconditional branch, strengths 0.825^(12-i))", "guess", (32, 48), 1, 20),
                 ("configs[4] VAE decode 512x512 batch 16", "vae", (64, 64), 16, 0),
                 ("VAE decode 256x384 batch 1 (the configs[1] image)", "vae", (32, 48), 1, 0)]
    out = []
    for name, kind, hw, batch, steps in jobs:
        try:
            if kind in ("denoise", "guess"):
                out.append(denoise_config(model, hw, batch, steps, 3, dist, dev, world, pk, name, guess=kind == "guess"))
            else:
                out.append(vae_config(model, hw, batch, dev, pk, name))
        except Exception as ex:  # noqa: BLE001
            out.append({"name": name, "error": f"{type(ex).__name__}: {ex}"[:300]})
            if dist is not None:
                raise  # ranks must not diverge around collectives
    return out


def run_gpu_arm(args, rank, world, local_rank):
    from stablediffusioneo_b200 import ops, synth
    from stablediffusioneo_b200.cldm.cldm import ControlLDM
    from stablediffusioneo_b200.cldm.ddim_hacked import DDIMSampler

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    pk = peaks()

    with torch.device(dev):
        model = ControlLDM().eval()
    synth.randomize_(model, seed=1234)
    for p in model.parameters():
        p.requires_grad_(False)
    sampler = DDIMSampler(model)

    x_T, ctx_c, ctx_u, hint = host_inputs(pin=True)
    cond = {"c_concat": [hint], "c_crossattn": [ctx_c]}
    uncond = {"c_concat": [hint], "c_crossattn": [ctx_u]}

    def sample_e2e():
        """Public API with host inputs: H2D of x_T / hint / contexts, 20 steps, D2H of the final latents."""
        samples, _ = sampler.sample(S_DDIM, 1, (4,) + LATENT_HW, cond, verbose=False, eta=0.0, x_T=x_T,
                                    unconditional_guidance_scale=CFG_SCALE, unconditional_conditioning=uncond)
        return samples.to("cpu", non_blocking=False)

    out = sample_e2e()  # builds the engine: packs weights, tunes the conv shapes, captures the step graph
    assert torch.isfinite(out).all() and out.abs().mean() > 1e-3
    sample_e2e()        # second image: the per-image prologue (hint encoder, cross-attention K/V) is captured too
    eng = sampler._engine
    launches_per_step = getattr(eng, "launches_per_step", None)

    # ---------------- value: K steps (x R repeats, >= 1 s), device-resident, CUDA events, max over ranks ----------------
    block_ms, repeats, eager_launches, clock_info = timed_steps(eng, args.steps, args.warmup, dist, dev, min_seconds=1.0,
                                                                clocks_for=local_rank if rank == 0 else None)
    elapsed_ms = block_ms
    value = world * args.steps / (elapsed_ms * 1e-3)

    # ---------------- e2e: public API, host buffers, >= 10 images ----------------
    n_img = max(10, math.ceil(args.steps / S_DDIM))
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    lat = []
    t0 = time.perf_counter()
    for _ in range(n_img):
        t1 = time.perf_counter()
        sample_e2e()   # returns after the D2H copy of the latents (synchronous)
        lat.append(time.perf_counter() - t1)
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
        # the only collective on the path: gather the ranks' final latents (NCCL all-gather), once, after the loop
        gathered = [torch.empty_like(eng.x_lat) for _ in range(world)]
        dist.all_gather(gathered, eng.x_lat)
    e2e_value = world * n_img * S_DDIM / e2e_s
    h2d = (x_T.numel() + ctx_c.numel() + ctx_u.numel() + 2 * hint.numel()) * 4   # hint is uploaded for cond and uncond
    d2h = x_T.numel() * 4
    lat.sort()

    # ---------------- full image (sample + VAE decode + uint8 image to host): p50 latency ----------------
    img_lat = []
    for _ in range(5):
        t1 = time.perf_counter()
        samples, _ = sampler.sample(S_DDIM, 1, (4,) + LATENT_HW, cond, verbose=False, eta=0.0, x_T=x_T,
                                    unconditional_guidance_scale=CFG_SCALE, unconditional_conditioning=uncond)
        u8 = model.decode_first_stage_u8(samples).cpu()
        img_lat.append(time.perf_counter() - t1)
    img_lat.sort()
    assert u8.shape == (1, 256, 384, 3)

    roof = conv_roofline(eng, pk) if rank == 0 else None
    extra = [] if args.no_other_configs else other_configs(model, args, dist, dev, world, pk, rank)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    step_ms = elapsed_ms / args.steps
    cfg = main_config(world)
    cfg["cuda_graph"] = eng.graph is not None
    line = {
        "metric": "denoise_steps_per_s", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": cfg,
        "timed_repeats": repeats, "timed_steps": repeats * args.steps,
        "images_per_s": value / S_DDIM,
        "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": h2d / S_DDIM, "d2h_bytes_per_step": d2h / S_DDIM,
                "images": n_img, "p50_denoise_latency_ms": 1000.0 * lat[len(lat) // 2],
                "api": "DDIMSampler.sample(S=20, x_T/hint/context on pinned host memory) -> latents on host"},
        "p50_image_latency_ms": 1000.0 * img_lat[len(img_lat) // 2],
        "gpu_launches": ((launches_per_step or 0) * args.steps * repeats + eager_launches) // repeats,
        "launches_per_step": launches_per_step,
        "clocks": clock_info,
        "roofline": roof,
        "step_roofline": {"tensor_frac": FLOPS_PER_STEP / (step_ms * 1e-3) / 1e12 / pk["bf16_sustained"],
                          "tflops": FLOPS_PER_STEP / (step_ms * 1e-3) / 1e12,
                          "weight_stream_frac": WEIGHT_BYTES_PER_STEP / (step_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                          "note": "whole step vs 740 GFLOP/step tensor bound and 2.442 GB/step weight-streaming HBM bound"},
        "other_configs": extra,
        "quality_gate": "compute_score PD (Inception-2048 features, compute_score.py:11-17) cannot run: pytorch_fid and its "
                        "pt_inception weights are in neither this image nor the GPU box (no network); the parity tests gate "
                        "the final uint8 image by PSNR >= 40 dB against the reference's image instead",
    }
    if world == 1 and not args.no_cpu_baseline:
        n, dt, cores = cpu_reference_steps(2, budget_s=25.0)
        line["cpu_baseline"] = {"value": n / dt, "unit": "steps/s", "cores": cores, "kind": "port",
                                "sample": f"{n} DDIM step(s) of the same workload on the host CPU, fp32 oracle port "
                                          f"of the reference modules, after 1 warm-up step"}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="sdeo", choices=["sdeo", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the other BASELINE configs (quick runs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the sdeo arm has no CPU fallback (use --impl reference)")
    run_gpu_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
