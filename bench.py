#!/usr/bin/env python
"""Headline benchmark: images/sec of the SD-v1.5 denoising loop on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config NAME]

Default workload = BASELINE configs[1] (``dpm_solver``): SD-v1.5 UNet, DPM-Solver++ (2M) 25 steps, 512x512
(latent 64x64), batch 16 per GPU, classifier-free guidance 7.5 (UNet batch 32), bf16, random-init weights,
synthetic prompt embeddings.  One bench "step" = one pass of the hot path over one batch.  The other BASELINE
configurations are selected with ``--config`` (the driver's contract run never passes it):

    deep_cache          configs[2]: DDIM 50 steps + DeepCache interval 3, global batch 64 sharded over the ranks
    consistency_model   configs[3]: LCM 4 steps, guidance 0 (no CFG), global batch 128 sharded over the ranks
    two_schedulers      configs[4]: DDIM -> DPM-Solver++ switch (N1 = 20, k = 10), 125 prompts per rank (1 000 on
                        8 GPUs) -> native VAE decode -> uint8 -> native CLIP towers -> NCCL all-gather of the
                        uint8 images and CLIP features -> CLIP score, collective INSIDE the timed e2e region

Own arm: ONE JSON line with
  value      images/s, whole job, inputs (latents, prompt embeddings) resident in HBM, loop only
             (the reference's time_metric scope, /root/reference/src/models.py:208,284-285)
  e2e        the same through the plugin call with PINNED HOST inputs copied H2D and the result copied D2H inside
             the timed region (two_schedulers: prompts as strings, images + features gathered, score read back)
  e2e_with_vae   prompts as strings -> tokenise -> text tower -> loop -> VAE decode -> uint8 images on the host
  roofline   dominant kernel (conv_gemm_kernel: every conv / linear of the UNet) vs the measured dense-bf16
             tensor peak, from per-operator CUDA-event times of the same plan; roofline_hbm: the HBM-bound
             kernels (GroupNorm, LayerNorm, fused latent update) vs the measured copy bandwidth
  gpu_library_baseline   the oracle UNet through stock PyTorch eager (cuDNN / cuBLASLt / flash-SDPA, bf16) on
             the same GPU at the same UNet batch -- the "existing Blackwell kernels" bar (rank 0, N=1 only)
  cpu_baseline  the oracle (CPU restatement of the reference path) on this box's host cores, bounded sample
Reference arm (--impl reference): the oracle's CPU path (there is no runnable reference: its arithmetic lives in
diffusers, absent here) on all host threads; each step is a bounded sample -- the first n of one image's 25 CFG
denoising steps at batch 1, fp32, n sized from a calibration step (25 = the whole image on a fast host), extrapolated
x25/n otherwise -- and the line's ``config`` says exactly which.
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

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GUIDANCE = 7.5
FLOP_PER_SAMPLE_FWD = 803.27e9

WORKLOADS = {
    "dpm_solver": dict(
        text="configs/dpm_solver_config.yaml: SD-v1.5 UNet, DPM-Solver++(2M) 25 steps, 512x512, batch 16 per GPU, "
             "CFG 7.5 (UNet batch 32), bf16",
        metric="images/sec (512x512, 25 steps DPM-Solver++, CFG 7.5)", scaling="weak", per_gpu=lambda w: 16,
        steps=25, guidance=7.5),
    "deep_cache": dict(
        text="configs/deep_cache_config.yaml: SD-v1.5 UNet, DDIM 50 steps + DeepCache interval 3 (branch 0), 512x512, "
             "global batch 64 sharded over the GPUs, CFG 7.5, bf16",
        metric="images/sec (512x512, DDIM 50 steps + DeepCache interval 3, CFG 7.5)", scaling="strong",
        per_gpu=lambda w: 64 // w, steps=50, guidance=7.5),
    "consistency_model": dict(
        text="configs/consistency_model_config.yaml: latent consistency model (LCM scheduler) 4 steps, 512x512, "
             "global batch 128 sharded over the GPUs, guidance 0 (no CFG), bf16",
        metric="images/sec (512x512, LCM 4 steps, no CFG)", scaling="strong", per_gpu=lambda w: 128 // w, steps=4,
        guidance=0.0),
    "two_schedulers": dict(
        text="configs/two_schedulers_config.yaml: DDIM -> DPM-Solver++ switch (N1=20, k=10: 21 UNet calls), 512x512, "
             "125 prompts per GPU in batches of <=32 (1 000 prompts on 8 GPUs), CFG 7.5, bf16; e2e adds VAE decode, "
             "uint8 quantise, CLIP towers, NCCL all-gather of images + features, CLIP score",
        metric="images/sec (512x512, DDIM->DPM-Solver++ 21 UNet calls, CFG 7.5)", scaling="weak",
        per_gpu=lambda w: 125, steps=21, guidance=7.5),
}


def _traffic():
    """DRAM bytes per conv_gemm_kernel launch (mean over the launches of one UNet step) from the committed ncu
    launch list (profiles/traffic.json, written by the round's profiling pass); None if absent."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(path))["conv_gemm_kernel"]
        return round((t["dram_read_bytes"] + t["dram_write_bytes"]) / t["launches"])
    except (OSError, KeyError, ValueError):
        return None


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d["bf16_tflops"], d["bf16_tflops_sustained"], d["hbm_gbs"], "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------- model construction
def build_model(device, workload, seed=29):
    import warnings

    from sonicdiffusionbayeslab_b200 import models as M
    from sonicdiffusionbayeslab_b200 import schedulers as S
    from sonicdiffusionbayeslab_b200.deepcache import DeepCacheSDHelper

    cfg = M.SD15_SCHEDULER_CONFIG
    cls = M.StableDiffusionModelTwoSchedulers if workload == "two_schedulers" else M.StableDiffusionModel
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")          # random-init weights ARE the benchmark's declared data
        model = cls.from_pretrained("runwayml/stable-diffusion-v1-5", torch_dtype=torch.bfloat16, seed=seed)
    model.to(device)
    dpmpp = dict(solver_order=2, algorithm_type="dpmsolver++", final_sigmas_type="zero")
    if workload == "dpm_solver":
        model.scheduler = S.DPMSolverScheduler.from_config(cfg, **dpmpp)
    elif workload == "deep_cache":
        model.scheduler = S.DDIMSchedulerMy.from_config(cfg)
        h = DeepCacheSDHelper(pipe=model)
        h.set_params(cache_interval=3, cache_branch_id=0)
        h.enable()
    elif workload == "consistency_model":
        model.scheduler = S.LCMScheduler.from_config(cfg)
    else:
        model.scheduler_first = S.DDIMSchedulerMy.from_config(cfg)
        model.scheduler_second = S.DPMSolverScheduler.from_config(cfg, **dpmpp)
    return model


def call_kwargs(workload):
    w = WORKLOADS[workload]
    if workload == "two_schedulers":
        return dict(num_inference_steps_first=20, num_inference_steps_second=20, num_step_switch=10,
                    guidance_scale=w["guidance"])
    return dict(num_inference_steps=w["steps"], guidance_scale=w["guidance"])


def _events():
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


# ----------------------------------------------------------------------------------------- same-box library baseline
def gpu_library_baseline(dev, unet_batch, engine_step_ms):
    """The oracle UNet in bf16 through stock PyTorch eager on the SAME GPU (cuDNN / cuBLASLt / flash SDPA, TF32 off)
    at the benchmarked UNet batch, plus four operator classes at the 64x64 level, each with this repo's kernel
    time for the same shape beside it (from the per-operator plan profile)."""
    import torch.nn.functional as F

    from oracle.unet import make_unet

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    net = make_unet(29).to(torch.bfloat16).to(dev).to(memory_format=torch.channels_last)
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(unet_batch, 4, 64, 64, device=dev, generator=g).bfloat16()
    ctx = torch.randn(unet_batch, 77, 768, device=dev, generator=g).bfloat16()
    t = torch.tensor(481, device=dev)

    def timed(fn, reps, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = _events()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    with torch.no_grad():
        unet_ms = timed(lambda: net(x, t, encoder_hidden_states=ctx), 5)
    del net
    torch.cuda.empty_cache()
    n = unet_batch
    ops = {}
    from sonicdiffusionbayeslab_b200 import kernels as K

    rows = n * 4096
    with torch.no_grad():
        # 3x3 convolution 320 -> 320 at 64x64
        xc = torch.randn(n, 320, 64, 64, device=dev).bfloat16().to(memory_format=torch.channels_last)
        w = torch.randn(320, 320, 3, 3, device=dev).bfloat16().to(memory_format=torch.channels_last)
        x_nhwc = xc.permute(0, 2, 3, 1).contiguous()
        wp, bias = K.pack_conv3x3_weight(w), torch.zeros(320, device=dev)
        out = torch.empty(rows, 320, device=dev, dtype=torch.bfloat16)
        fl = 2 * rows * 320 * 2880
        a = timed(lambda: F.conv2d(xc, w, padding=1), 20)
        b = timed(lambda: K.conv_gemm(x_nhwc, wp, 320, taps=9, n_img=n, H=64, W=64, bias=bias, out=out), 20)
        ops["conv3x3 320->320 @64x64"] = {"torch_us": round(a * 1e3, 1), "torch_tflops": round(fl / a / 1e9, 1),
                                         "sonic_us": round(b * 1e3, 1), "sonic_tflops": round(fl / b / 1e9, 1)}
        # self-attention, 4096 tokens, 8 heads of 40
        q = torch.randn(n, 8, 4096, 40, device=dev).bfloat16()
        qkv = torch.randn(rows, 960, device=dev).bfloat16()
        ao = torch.empty(rows, 320, device=dev, dtype=torch.bfloat16)
        fl = 4 * n * 8 * 4096 * 4096 * 40
        a = timed(lambda: F.scaled_dot_product_attention(q, q, q), 10)
        b = timed(lambda: K.attention(qkv[:, :320], qkv[:, 320:640], qkv[:, 640:], batch=n, heads=8, seq_q=4096,
                                      seq_k=4096, head_dim=40, out=ao), 10)
        ops["self-attention 4096 tokens, 8 heads, d=40"] = {
            "torch_us": round(a * 1e3, 1), "torch_tflops": round(fl / a / 1e9, 1),
            "sonic_us": round(b * 1e3, 1), "sonic_tflops": round(fl / b / 1e9, 1)}
        # GroupNorm(32) + SiLU
        gw, gb = torch.ones(320, device=dev), torch.zeros(320, device=dev)
        g16, b16 = gw.bfloat16(), gb.bfloat16()
        by = 2 * xc.numel() * 2
        x2d = x_nhwc.view(rows, 320)
        a = timed(lambda: F.silu(F.group_norm(xc, 32, g16, b16, 1e-5)), 20)
        b = timed(lambda: K.groupnorm(x2d, gw, gb, n_img=n, hw=4096, out=out), 20)
        ops["GroupNorm(32)+SiLU 320ch @64x64"] = {"torch_us": round(a * 1e3, 1), "torch_gbs": round(by / a / 1e6, 1),
                                                 "sonic_us": round(b * 1e3, 1), "sonic_gbs": round(by / b / 1e6, 1),
                                                 "note": "sonic = statistics pass + apply (in the UNet the statistics "
                                                         "come from the producing GEMM's epilogue)"}
        # LayerNorm (standalone kernel; in the UNet it is folded into the surrounding GEMMs)
        a = timed(lambda: F.layer_norm(x2d, (320,), g16, b16, 1e-5), 20)
        b = timed(lambda: K.layernorm(x2d, gw, gb, out=out), 20)
        ops["LayerNorm 320 over 64x64 tokens"] = {"torch_us": round(a * 1e3, 1), "torch_gbs": round(by / a / 1e6, 1),
                                                 "sonic_us": round(b * 1e3, 1), "sonic_gbs": round(by / b / 1e6, 1)}
    for v in ops.values():
        v["speedup"] = round(v["torch_us"] / v["sonic_us"], 2)
    ops["_how"] = "every operator timed ALONE, back to back (3 warm-up + 10-20 timed launches, CUDA events), same tensors"
    return {"what": f"oracle UNet (same architecture / weights layout) in bf16 through stock PyTorch "
                    f"{torch.__version__} eager, channels_last, cudnn.benchmark, TF32 off, UNet batch {unet_batch}",
            "unet_step_ms": round(unet_ms, 2), "sonic_unet_step_ms": round(engine_step_ms, 2),
            "speedup": round(unet_ms / engine_step_ms, 2), "ops": ops}


# ----------------------------------------------------------------------------------------- own arm
def run_own(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU path (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")               # these two levels print a version banner on STDOUT, where only
                                                       # the JSON line belongs; INFO (what a harness may set) is left alone
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)

    wl = args.config
    W = WORKLOADS[wl]
    B = W["per_gpu"](world)
    if B < 1:
        raise SystemExit(f"bench.py: workload {wl} does not shard over {world} GPUs")
    model = build_model(dev, wl)
    kw = call_kwargs(wl)
    if wl == "two_schedulers":
        line = run_two_schedulers(args, model, kw, B, rank, world, local, dev, dist)
    else:
        line = run_loop_workload(args, model, kw, wl, B, rank, world, local, dev, dist)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def _barrier(dist):
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()


def run_loop_workload(args, model, kw, wl, B, rank, world, local, dev, dist):
    W = WORKLOADS[wl]
    # Strong-scaling workloads (one GLOBAL batch sharded over the ranks) draw the global tensors and keep this
    # rank's rows -- the single-GPU noise (rng_rows); the weak one gives every rank its own seeded batch.
    g = torch.Generator(device="cpu").manual_seed(29 + (rank if W["scaling"] == "weak" else 0))
    total = B * world if W["scaling"] == "strong" else B
    lo = rank * B if W["scaling"] == "strong" else 0
    pe_host = torch.randn(total, 77, 768, generator=g).to(torch.bfloat16)[lo:lo + B].contiguous().pin_memory()
    ne_host = torch.randn(total, 77, 768, generator=g).to(torch.bfloat16)[lo:lo + B].contiguous().pin_memory()
    lat_host = torch.randn(total, 4, 64, 64, generator=g).to(torch.bfloat16)[lo:lo + B].contiguous().pin_memory()
    out_host = torch.empty(B, 4, 64, 64, dtype=torch.bfloat16).pin_memory()
    pe, ne, lat = pe_host.to(dev), ne_host.to(dev), lat_host.to(dev)
    cfg_on = W["guidance"] > 1
    noise_gen = torch.Generator(device=dev)          # on the device, like base_experiment.py:51-53
    rows = (lo, lo + B, total) if W["scaling"] == "strong" else None

    def extra():
        if wl != "consistency_model":
            return {}
        noise_gen.manual_seed(7)                      # LCM draws fresh noise every step from this generator
        return {"generator": noise_gen, "rng_rows": rows}

    def step_resident():
        _, secs, _ = model(prompt_embeds=pe, negative_prompt_embeds=ne if cfg_on else None, latents=lat,
                           output_type="latent", **kw, **extra())
        return secs

    def step_e2e():
        o, _, _ = model(prompt_embeds=pe_host.to(dev, non_blocking=True),
                        negative_prompt_embeds=ne_host.to(dev, non_blocking=True) if cfg_on else None,
                        latents=lat_host.to(dev, non_blocking=True), output_type="latent", **kw, **extra())
        out_host.copy_(o.images, non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_resident()
    step_e2e()

    # ---- timed region 1: inputs resident in HBM, device-timed (CUDA events), max over ranks
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _barrier(dist)
    ev0, ev1 = _events()
    ev0.record()
    for _ in range(args.steps):
        step_resident()
    ev1.record()
    _barrier(dist)
    ms_total = ev0.elapsed_time(ev1)
    # ---- timed region 2: end to end through the plugin call with pinned host buffers
    _barrier(dist)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    _barrier(dist)
    e2e_s = time.perf_counter() - t0
    # ---- timed region 3 (default workload): strings in, uint8 images out -- text tower + loop + VAE decode
    vae_s = None
    if wl == "dpm_solver":
        from sonicdiffusionbayeslab_b200.dataset.dataset import synthetic_prompts

        prompts = synthetic_prompts(B, seed=29 + rank)
        img_host = torch.empty(B, 3, 512, 512, dtype=torch.uint8).pin_memory()
        model.decode_x0_preds = False                 # per-step x0 previews (models.py:295-302) are not decoded

        def step_vae():
            o, _, _ = model(prompts, latents=lat_host.to(dev, non_blocking=True), output_type="pt", **kw)
            img_host.copy_((o.images * 255).to(torch.uint8), non_blocking=True)
            torch.cuda.synchronize()

        step_vae()
        _barrier(dist)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_vae()
        _barrier(dist)
        vae_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms_total, e2e_s, vae_s or 0.0], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_s, vae_s = t.tolist()
        vae_s = vae_s or None

    images = B * args.steps * world
    value = images / (ms_total / 1e3)
    if rank != 0:
        return None
    # ---- rooflines from per-operator CUDA-event times of the recorded plan (rank 0)
    import ctypes

    eng = model.engine(B, cfg_on)
    side = torch.cuda.Stream(device=dev)                    # kernels are timed on the stream they run on
    with torch.cuda.stream(side):
        eng.plans["full"].profile(ctypes.c_void_p(side.cuda_stream))          # warm
        prof = eng.plans["full"].profile(ctypes.c_void_p(side.cuda_stream))
    torch.cuda.synchronize()
    log = eng.plans["full"].log
    names = {0: "conv_gemm_kernel", 1: "attention kernels", 2: "groupnorm (gn_cluster_kernel)", 3: "layernorm_kernel",
             4: "layout kernels", 5: "timestep gemv"}
    agg, hbm, ln_side = {}, {2: [0.0, 0.0, 0], 3: [0.0, 0.0, 0]}, [0.0, 0.0, 0]
    for (kind, ms, fl), text in zip(prof, log):
        a = agg.setdefault(kind, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += ms
        a[2] += fl
        if kind in hbm:                                     # "groupnorm rows=R C=C ..." / "layernorm rows=R C=C"
            f = dict(p.split("=") for p in text.split()[1:] if "=" in p)
            if text.startswith("ln_side"):                  # folded LayerNorm: partials in, 16 B of side row + rstd out
                ln_side[0] += int(f["rows"]) * (int(f["parts"]) * 8 + 20)
                ln_side[1] += ms
                ln_side[2] += 1
                continue
            hbm[kind][0] += 2.0 * int(f["rows"]) * int(f["C"]) * 2       # read once + write once, bf16
            hbm[kind][1] += ms
            hbm[kind][2] += 1
    tot_ms = sum(a[1] for a in agg.values())
    burst, sustained, hbm_peak, how = _peaks()
    n_unet = len(model.last_step_kinds) or W["steps"]
    n_full = n_unet - model.last_step_kinds.count("cached")
    # The timed region replays the plan as a CUDA graph; the per-operator pass is eager (event pairs around every
    # launch) and runs later, at whatever clocks the power cap then allows.  Each operator class is therefore given
    # its SHARE of the eager pass applied to the graph-replay time of a full UNet step.
    full_step_ms = full_unet_step_ms(model, eng, dev)
    scale = full_step_ms / tot_ms
    kernels = [{"kernel": names[k], "ops": c, "ms": round(ms * scale, 3), "share": round(ms / tot_ms, 4),
                "tflops": round(fl / (ms * scale) / 1e9, 1) if fl else None}
               for k, (c, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    gemm = agg[0]
    achieved = gemm[2] / (gemm[1] * scale) / 1e9
    n_launch, plan_flops = eng.stats("full")
    unet_ms = ms_total / args.steps / n_unet
    # fused latent update: one launch per step; timed alone (CUDA events around 200 launches at this batch)
    upd = time_latent_update(dev, B, cfg_on)
    roof_hbm = {}
    for kind, label in ((2, "groupnorm"), (3, "layernorm")):
        by, ms, n_ops = hbm[kind]
        ms *= scale
        if ms > 0:
            roof_hbm[label] = {"bound": "hbm", "achieved": round(by / ms / 1e6, 1), "peak": hbm_peak, "unit": "GB/s",
                               "frac": round(by / ms / 1e6 / hbm_peak, 4), "ops": n_ops,
                               "bytes_per_step": int(by), "ms_per_step": round(ms, 3)}
        else:
            roof_hbm[label] = {"ops": 0, "note": "no such kernel in the plan"}
    if ln_side[2]:
        roof_hbm["layernorm"] = {
            "ops": 0, "folded_into_gemms": ln_side[2],
            "note": "the LayerNorms make no pass over the activations: row statistics come from the producer GEMM's "
                    "epilogue, ln_side_kernel turns them into a 16-byte side row + rstd, the consumer GEMM applies them",
            "ln_side_kernel": {"launches": ln_side[2], "ms_per_step": round(ln_side[1] * scale, 3),
                               "bytes_per_step": int(ln_side[0]), "bound": "latency (a few MB per launch)"}}
    roof_hbm["latent_update_kernel"] = upd | {"peak": hbm_peak, "frac": round(upd["achieved"] / hbm_peak, 4)}
    n_cached = n_unet - n_full
    launches_per_image_loop = (n_unet - n_cached) * (n_launch + 1) + n_cached * (
        (eng.stats("cached")[0] if "cached" in eng.plans else 0) + 1)
    line = {
        "metric": W["metric"], "value": round(value, 3), "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms_total / args.steps, 3),
        "higher_is_better": True, "scaling": W["scaling"], "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic (random-init SD-v1.5 weights, N(0,1) prompt embeddings and latents, seed 29)",
        "config": {"workload": W["text"], "per_gpu_batch": B, "global_batch": B * world,
                   "l2": "working set larger than L2 (weights 1.7 GB + streamed activations)"},
        "e2e": {"value": round(images / e2e_s, 3), "unit": "images/s",
                "h2d_bytes_per_step": int(pe_host.numel() * 2 * (2 if cfg_on else 1) + lat_host.numel() * 2),
                "d2h_bytes_per_step": int(out_host.numel() * 2),
                # its OWN timed region (host wall clock around K plugin calls incl. the copies, max over ranks)
                "ms_per_step": round(e2e_s / args.steps * 1e3, 3), "timer": "perf_counter around synchronised calls"},
        "gpu_launches": int(args.steps * 2 * (eng.stats("ctx")[0] + launches_per_image_loop)),
        "unet_step_ms": round(unet_ms, 3),
        "unet_calls_per_image": n_unet, "deepcache_cached_steps": n_cached or None,
        "roofline": {"kernel": "conv_gemm_kernel (all conv3x3/conv1x1/linear launches of one UNet forward)",
                     "bound": "tensor", "achieved": round(achieved, 1), "peak": sustained, "unit": "TFLOP/s",
                     "frac": round(achieved / sustained, 4), "peak_kind": f"bf16_tflops_sustained ({how})",
                     "frac_of_burst": round(achieved / burst, 4), "traffic": _traffic(),
                     "algorithmic_per_launch": round(gemm[2] / gemm[0]), "launches_per_step": gemm[0],
                     "full_unet_step_ms": round(full_step_ms, 3), "eager_profile_total_ms": round(tot_ms, 3),
                     "note": "achieved = sum of 2*M*N*K over the step's conv/linear launches / (their share of the "
                             "per-operator CUDA-event pass x the graph-replay time of one full UNet step, both on the "
                             "launching stream); traffic = mean DRAM bytes per launch from the committed ncu launch "
                             "list (cold cache per launch)"},
        "roofline_hbm": roof_hbm,
        "kernels": kernels,
        "clocks": clocks,
    }
    if wl == "dpm_solver":
        line["unet_tflops"] = round(2 * B * FLOP_PER_SAMPLE_FWD / unet_ms / 1e9, 1)
    if vae_s:
        line["e2e_with_vae"] = {
            "value": round(images / vae_s, 3), "unit": "images/s",
            "includes": "hash tokenise + native CLIP-L text tower (prompt and empty negative prompt) + denoising loop "
                        "+ native VAE decode + uint8 quantise + D2H of the images; x0 previews not decoded",
            "h2d_bytes_per_step": int(lat_host.numel() * 2 + 2 * B * 77 * 8),
            "d2h_bytes_per_step": int(B * 3 * 512 * 512)}
    if world == 1 and wl == "dpm_solver" and not args.no_library_baseline:
        try:
            line["gpu_library_baseline"] = gpu_library_baseline(dev, 2 * B if cfg_on else B, unet_ms)
        except Exception as e:                              # noqa: BLE001  (an OOM here must not lose the bench line)
            line["gpu_library_baseline"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline()
    return line


def full_unet_step_ms(model, eng, dev, reps=10):
    """Graph-replay time of ONE full UNet forward (what the timed region runs), CUDA events, back to back."""
    for _ in range(3):
        eng.forward(481.0)
    a, b = _events()
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        eng.forward(481.0)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def time_latent_update(dev, B, cfg_on):
    """The fused CFG + DPM-Solver++ 2nd-order update alone: a CUDA graph of 100 launches, replayed 5 times between
    CUDA events (eager Python calls would time the host, ~100 us of coefficient arithmetic per step, which the real
    loop hides behind the UNet graph)."""
    from sonicdiffusionbayeslab_b200 import kernels as K

    x = torch.randn(B, 4, 64, 64, device=dev).bfloat16()
    eu, et, m1 = torch.randn_like(x), torch.randn_like(x), torch.randn_like(x)
    m0, x0 = torch.empty_like(x), torch.empty_like(x[:1])
    c = dict(guidance=GUIDANCE, m_x=1.1, m_e=-0.5, x0_x=1.1, x0_e=-0.5, c_x=0.9, c_m0=0.3, c_h1=-0.1)
    n = 100

    def launch():
        K.latent_update(c, eu, x, eps_text=et if cfg_on else None, h1=m1, out_sample=x, out_m0=m0, out_x0=x0)

    launch()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(n):
                launch()
    torch.cuda.synchronize()
    g.replay()
    a, b = _events()
    torch.cuda.synchronize()
    a.record()
    for _ in range(5):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) / (5 * n) * 1e3
    per_image = (6 if cfg_on else 5) * 32768                # eps_u, eps_c, x, m1 read; x', m0 written (+ x0 of image 0)
    by = per_image * B + 32768
    return {"bound": "hbm (latency-bound at this size)", "achieved": round(by / us / 1e3, 1), "unit": "GB/s",
            "us_per_launch": round(us, 2), "bytes_per_launch": by,
            "note": f"{by} algorithmic bytes per launch ({B} latents): smaller than one wave's worth of traffic, so "
                    "launch + DRAM latency, not bandwidth, set the time; one launch per denoising step"}


def run_two_schedulers(args, model, kw, n_prompts, rank, world, local, dev, dist):
    """BASELINE configs[4]: each rank generates its contiguous block of the global prompt list (125 prompts,
    batches of <= 32), decodes, scores with the CLIP towers, and the uint8 images + features are all-gathered."""
    from sonicdiffusionbayeslab_b200 import dist as D
    from sonicdiffusionbayeslab_b200.dataset.dataset import synthetic_prompts
    from sonicdiffusionbayeslab_b200.metrics.metrics import ClipScoreMetric
    import warnings

    W = WORKLOADS["two_schedulers"]
    prompts_all = synthetic_prompts(n_prompts * world, seed=29)
    mine = prompts_all[rank * n_prompts:(rank + 1) * n_prompts]
    batches = [mine[i:i + 32] for i in range(0, len(mine), 32)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        clip = ClipScoreMetric().to(dev)
    model.decode_x0_preds = False
    gen = torch.Generator(device="cpu")
    state = {}

    def job(loop_only):
        """One pass over this rank's prompts.  Returns (loop seconds, score or None)."""
        gen.manual_seed(29)
        loop_s, imgs, fi, ft = 0.0, [], [], []
        done = 0
        for b in batches:
            # the shared generator is replayed over the batches of lower ranks once, then consumed in order
            if done == 0:
                for r in range(rank):
                    for i in range(0, n_prompts, 32):
                        model([""] * min(32, n_prompts - i), generator=gen, output_type="pt", rng_only=True, **kw)
            o, secs, _ = model(b, generator=gen, output_type="latent" if loop_only else "pt", **kw)
            loop_s += secs
            done += len(b)
            if not loop_only:
                u8 = D.quantise_uint8(o.images.float())
                a, t_ = clip.features(u8, b)
                imgs.append(u8)
                fi.append(a)
                ft.append(t_)
        if loop_only:
            return loop_s, None
        gi, gf, gt = D.gather_images_and_features(torch.cat(imgs), torch.cat(fi), torch.cat(ft))
        score = D.clip_score_from_features(gf, gt)
        state["gathered"] = (gi.shape[0], gi.numel() + (gf.numel() + gt.numel()) * 4)
        return loop_s, float(score.item())               # D2H of the metric

    warm = max(args.warmup, 3)
    for i in range(warm):
        job(loop_only=i < warm - 1)                       # the last warm-up pass exercises decode + CLIP + gather
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _barrier(dist)
    ev0, ev1 = _events()
    ev0.record()
    loop_s = sum(job(True)[0] for _ in range(args.steps))
    ev1.record()
    _barrier(dist)
    ms_total = ev0.elapsed_time(ev1)
    _barrier(dist)
    t0 = time.perf_counter()
    score = None
    for _ in range(args.steps):
        _, score = job(False)
    _barrier(dist)
    e2e_s = time.perf_counter() - t0
    # the collective alone, for its share of the e2e region
    n_img, gathered = state["gathered"]
    u8 = torch.zeros(n_prompts, 3, 512, 512, dtype=torch.uint8, device=dev)
    f = torch.zeros(n_prompts, 512, device=dev)
    _barrier(dist)
    a, b = _events()
    a.record()
    D.gather_images_and_features(u8, f, f)
    b.record()
    _barrier(dist)
    gather_ms = a.elapsed_time(b)
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms_total, e2e_s, gather_ms, loop_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_s, gather_ms, loop_s = t.tolist()
    if rank != 0:
        return None
    images = n_prompts * world * args.steps
    unet_launches = 0
    for b in batches:
        eng = model.engine(len(b), True)
        unet_launches += eng.stats("ctx")[0] + W["steps"] * (eng.stats("full")[0] + 1)
    return {
        "metric": W["metric"], "value": round(images / (ms_total / 1e3), 3), "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": warm, "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True,
        "scaling": W["scaling"], "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic (random-init SD-v1.5 / VAE / CLIP weights, seeded synthetic captions, seed 29)",
        "config": {"workload": W["text"], "per_gpu_prompts": n_prompts, "global_prompts": n_prompts * world,
                   "l2": "working set larger than L2 (weights 1.7 GB + streamed activations)"},
        "e2e": {"value": round(images / e2e_s, 3), "unit": "images/s",
                "h2d_bytes_per_step": int(n_prompts * 2 * 77 * 8),
                "d2h_bytes_per_step": 4,
                "includes": "hash tokenise + text tower + loop + VAE decode + uint8 + CLIP towers + all-gather + score"},
        "loop_seconds_per_step": round(loop_s / args.steps, 3),
        "collective": {"op": "all_gather (uint8 images + fp32 image/text features)", "backend": "nccl" if world > 1 else "none",
                       "bytes_gathered_per_rank": int(gathered), "images_gathered": int(n_img),
                       "ms": round(gather_ms, 3), "share_of_e2e": round(gather_ms / (e2e_s / args.steps * 1e3), 5)},
        "clip_score": score,
        "gpu_launches": int(2 * args.steps * unet_launches),     # UNet plans + fused updates of both timed regions;
        "clocks": clocks,                                        # the text / VAE / CLIP plans of e2e come on top
    }


# ----------------------------------------------------------------------------------------- CPU oracle arms
_ORACLE = {}


def _oracle_run(n_steps, threads):
    """Seconds the oracle (reference restatement) takes on the CPU for the first ``n_steps`` steps of the 25-step
    DPM-Solver++ trajectory at batch 1 with CFG 7.5 (2 UNet sample-forwards per step), fp32, ``threads`` threads."""
    from oracle.pipeline import denoise
    from oracle.schedulers import SD15_SCHEDULER_CONFIG, DPMSolverScheduler
    from oracle.unet import make_unet

    torch.set_num_threads(threads)
    if "net" not in _ORACLE:
        g = torch.Generator().manual_seed(29)
        _ORACLE.update(net=make_unet(29), pe=torch.randn(1, 77, 768, generator=g),
                       ne=torch.randn(1, 77, 768, generator=g), lat=torch.randn(1, 4, 64, 64, generator=g))

    class Trunc(DPMSolverScheduler):            # run only the first n_steps of the 25-step schedule
        def set_timesteps(self, *a, **k):
            super().set_timesteps(*a, **k)
            self.timesteps = self.timesteps[:n_steps]

    sched = Trunc.from_config(SD15_SCHEDULER_CONFIG, solver_order=2, algorithm_type="dpmsolver++",
                              final_sigmas_type="zero")
    t0 = time.perf_counter()
    denoise(_ORACLE["net"], sched, _ORACLE["pe"], _ORACLE["ne"], _ORACLE["lat"], STEPS_PER_IMAGE,
            guidance_scale=GUIDANCE)
    return time.perf_counter() - t0


STEPS_PER_IMAGE = 25


def _sample_size(step_seconds, budget_s):
    """Denoising steps per timed CPU sample: as many of the image's 25 as fit ``budget_s`` (25 = a whole image, no
    extrapolation), at least one."""
    return max(1, min(STEPS_PER_IMAGE, int(budget_s / max(step_seconds, 1e-3))))


def _sample_text(n, s):
    whole = n == STEPS_PER_IMAGE
    return (f"{'the whole 25-step image' if whole else f'the first {n} of the 25 denoising steps of one image'} at "
            f"batch 1, CFG 7.5 (2 UNet sample-forwards per step), fp32, {s:.2f} s/step"
            + ("" if whole else f", extrapolated x{STEPS_PER_IMAGE}/{n}"))


def cpu_baseline(budget_s=15.0):
    """The oracle on the host cores next to the GPU number: one warm-up step (which also sizes the sample), then ONE
    timed sample of as many denoising steps as fit ~``budget_s`` seconds."""
    threads = len(os.sched_getaffinity(0))
    _oracle_run(1, threads)                                    # allocations, thread pool
    n = _sample_size(_oracle_run(1, threads), budget_s)
    s = _oracle_run(n, threads) / n
    return {"value": round(1.0 / (s * STEPS_PER_IMAGE), 6), "unit": "images/s", "cores": threads, "kind": "port",
            "sample": _sample_text(n, s), "sampled_steps": n, "extrapolated": n != STEPS_PER_IMAGE,
            "gflops": round(2 * FLOP_PER_SAMPLE_FWD / s / 1e9, 1)}


def run_reference(args):
    """``--impl reference``: the reference's CPU path (its oracle restatement -- diffusers is not installable here) on
    all host threads.  A "step" is one bounded sample of the workload: the first n denoising steps of one image, n
    sized from a calibration step so that one sample takes at most ~27 s (n = 25, a whole image, at <= 1.08 s/step)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    if args.config != "dpm_solver":
        print(json.dumps({"impl": "reference", "unavailable": "the CPU oracle arm is defined for the default "
                                                               "workload (dpm_solver) only"}), flush=True)
        return
    threads = len(os.sched_getaffinity(0))
    _oracle_run(1, threads)                                    # warm-up (untimed) ...
    n = _sample_size(_oracle_run(1, threads), 27.0)            # ... and calibration
    times = [_oracle_run(n, threads) / n for _ in range(max(1, min(args.steps, 3)))]
    s = statistics.mean(times)
    value = 1.0 / (s * STEPS_PER_IMAGE)
    W = WORKLOADS["dpm_solver"]
    line = {
        "impl": "reference", "metric": W["metric"],
        "value": round(value, 6), "unit": "images/s", "n_gpus": int(os.environ.get("WORLD_SIZE", 1)),
        "steps": len(times), "warmup": 2, "ms_per_step": round(s * 1e3 * STEPS_PER_IMAGE, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # what THIS arm really ran -- not the GPU arm's batch-16 bf16 workload
        "config": {"workload": "bounded CPU sample of configs/dpm_solver_config.yaml: the oracle restatement of the "
                               f"reference loop, {_sample_text(n, s)}, all host threads; images/s = "
                               "1 / (25 x seconds per step)",
                   "per_gpu_batch": 1, "dtype": "f32", "extrapolated": n != STEPS_PER_IMAGE, "sampled_steps": n,
                   "samples": len(times), "steps_per_image": STEPS_PER_IMAGE, "gpu_arm_workload": W["text"]},
        "cpu_baseline": {"value": round(value, 6), "unit": "images/s", "cores": threads, "kind": "port",
                         "sample": "oracle restatement of the reference path (diffusers is not installable here), all "
                                   f"host threads: {len(times)} sample(s) of {_sample_text(n, s)}"},
        "e2e": {"value": round(value, 6), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--config", default="dpm_solver", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
