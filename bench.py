#!/usr/bin/env python
"""Headline benchmark: images/sec of the SD-v1.5 denoising loop on B200 (BASELINE.json metric).

Workload (configs[1]): SD-v1.5 UNet, DPM-Solver++ (2M) 25 steps, 512x512 (latent 64x64), batch 16,
classifier-free guidance 7.5 (UNet batch 32), bf16, random-init weights, synthetic prompt
embeddings.  One bench "step" = one pass of the hot path over one batch: 25 x (UNet plan replay +
fused CFG/scheduler update) for 16 images.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Own arm (default): prints ONE JSON line with
  value      images/s, whole job, inputs (latents, prompt embeddings) resident in HBM, loop only
             (the reference's time_metric scope, /root/reference/src/models.py:208,284-285)
  e2e        the same loop through the plugin call ``model(prompt_embeds=, latents=, ...)`` with
             PINNED HOST inputs copied H2D and the final latents copied D2H inside the timed region
  roofline   dominant kernel (conv_gemm_kernel: every conv / linear of the UNet) vs the measured
             dense-bf16 tensor peak, from per-operator CUDA-event times of the same plan
  cpu_baseline  the oracle (CPU restatement of the reference path) on this box's host cores, on a
             bounded sample (rank 0, N=1 only)
Reference arm (--impl reference): the oracle's CPU path (there is no runnable reference: its
arithmetic lives in diffusers, absent here) timed on all host threads; each step is a bounded
sample (one CFG denoising step at batch 1) extrapolated to the 25-step workload.
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

STEPS_PER_IMAGE = 25
BATCH = 16
GUIDANCE = 7.5
FLOP_PER_SAMPLE_FWD = 803.27e9
CONFIG = {"workload": "configs/dpm_solver_config.yaml: SD-v1.5 UNet, DPM-Solver++(2M) 25 steps, 512x512, "
                      "batch 16 per GPU, CFG 7.5 (UNet batch 32), bf16",
          "per_gpu_batch": BATCH, "l2": "working set larger than L2 (weights 1.7 GB + streamed activations)"}


def _traffic():
    """DRAM bytes per conv_gemm_kernel launch (mean over the 194 launches of one UNet step) from the committed ncu
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


def build_model(device, seed=29):
    from sonicdiffusionbayeslab_b200 import models as M
    from sonicdiffusionbayeslab_b200 import schedulers as S

    model = M.StableDiffusionModel.from_pretrained("runwayml/stable-diffusion-v1-5", torch_dtype=torch.bfloat16,
                                                   seed=seed)
    model.to(device)
    model.scheduler = S.DPMSolverScheduler.from_config(M.SD15_SCHEDULER_CONFIG, solver_order=2,
                                                       algorithm_type="dpmsolver++", final_sigmas_type="zero")
    return model


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

        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"          # NCCL prints its version banner on stdout otherwise

        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)

    model = build_model(dev)
    g = torch.Generator(device="cpu").manual_seed(29 + rank)
    pe_host = torch.randn(BATCH, 77, 768, generator=g).to(torch.bfloat16).pin_memory()
    ne_host = torch.randn(BATCH, 77, 768, generator=g).to(torch.bfloat16).pin_memory()
    lat_host = torch.randn(BATCH, 4, 64, 64, generator=g).to(torch.bfloat16).pin_memory()
    out_host = torch.empty(BATCH, 4, 64, 64, dtype=torch.bfloat16).pin_memory()
    pe, ne, lat = pe_host.to(dev), ne_host.to(dev), lat_host.to(dev)

    def step_resident():
        _, secs, _ = model(prompt_embeds=pe, negative_prompt_embeds=ne, latents=lat,
                           num_inference_steps=STEPS_PER_IMAGE, guidance_scale=GUIDANCE, output_type="latent")
        return secs

    def step_e2e():
        o, _, _ = model(prompt_embeds=pe_host.to(dev, non_blocking=True),
                        negative_prompt_embeds=ne_host.to(dev, non_blocking=True),
                        latents=lat_host.to(dev, non_blocking=True), num_inference_steps=STEPS_PER_IMAGE,
                        guidance_scale=GUIDANCE, output_type="latent")
        out_host.copy_(o.images, non_blocking=True)
        torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_resident()
    step_e2e()

    # ---- timed region 1: inputs resident in HBM, device-timed (CUDA events), max over ranks
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step_resident()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    # ---- timed region 2: end to end through the plugin call with pinned host buffers
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms_total, e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_s = t.tolist()

    images = BATCH * args.steps * world
    value = images / (ms_total / 1e3)
    e2e_value = images / e2e_s

    # ---- roofline of the dominant kernel from per-operator CUDA-event times (rank 0)
    line = None
    if rank == 0:
        eng = model.engine(BATCH, True)
        import ctypes

        side = torch.cuda.Stream(device=dev)                    # kernels are timed on the stream they run on
        with torch.cuda.stream(side):
            eng.plans["full"].profile(ctypes.c_void_p(side.cuda_stream))          # warm
            prof = eng.plans["full"].profile(ctypes.c_void_p(side.cuda_stream))
        torch.cuda.synchronize()
        names = {0: "conv_gemm_kernel", 1: "attention_kernel", 2: "groupnorm(stats+apply)", 3: "layernorm_kernel",
                 4: "layout kernels", 5: "timestep gemv"}
        agg = {}
        for kind, ms, fl in prof:
            a = agg.setdefault(kind, [0, 0.0, 0.0])
            a[0] += 1
            a[1] += ms
            a[2] += fl
        tot_ms = sum(a[1] for a in agg.values())
        burst, sustained, hbm, how = _peaks()
        kernels = []
        for kind, (cnt, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            kernels.append({"kernel": names[kind], "ops": cnt, "ms": round(ms, 3), "share": round(ms / tot_ms, 4),
                            "tflops": round(fl / ms / 1e9, 1) if fl else None})
        gemm = agg[0]
        achieved = gemm[2] / gemm[1] / 1e9                      # TFLOP/s over all conv/linear launches
        n_launch, plan_flops = eng.stats("full")
        unet_ms = ms_total / args.steps / STEPS_PER_IMAGE
        line = {
            "metric": "images/sec (512x512, 25 steps DPM-Solver++, CFG 7.5)",
            "value": round(value, 3), "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_total / args.steps, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic (random-init SD-v1.5 weights, N(0,1) prompt embeddings and latents, seed 29)",
            "config": dict(CONFIG),
            "e2e": {"value": round(e2e_value, 3), "unit": "images/s",
                    "h2d_bytes_per_step": int(pe_host.numel() * 2 * 2 + lat_host.numel() * 2),
                    "d2h_bytes_per_step": int(out_host.numel() * 2)},
            "gpu_launches": int(args.steps * 2 * (eng.stats("ctx")[0] + STEPS_PER_IMAGE * (n_launch + 1))),
            "unet_step_ms": round(unet_ms, 3),
            "unet_tflops": round(2 * BATCH * FLOP_PER_SAMPLE_FWD / unet_ms / 1e9, 1),
            "roofline": {"kernel": "conv_gemm_kernel (all conv3x3/conv1x1/linear launches of one UNet forward)",
                         "bound": "tensor", "achieved": round(achieved, 1), "peak": sustained, "unit": "TFLOP/s",
                         "frac": round(achieved / sustained, 4), "peak_kind": f"bf16_tflops_sustained ({how})",
                         "frac_of_burst": round(achieved / burst, 4), "traffic": _traffic(),
                         "algorithmic_per_launch": round(gemm[2] / gemm[0]), "launches_per_step": gemm[0],
                         "note": "achieved = sum of 2*M*N*K over the step's conv/linear launches / sum of their CUDA-event "
                                 "durations (same stream); traffic = mean DRAM bytes per launch from the committed ncu "
                                 "launch list (cold cache per launch)"},
            "kernels": kernels,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(sample_steps=1)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def _oracle_cfg_step_seconds(n_steps, threads):
    """Times the oracle (reference restatement) on CPU: batch 1, CFG 7.5, DPM-Solver++ steps."""
    from oracle.pipeline import denoise
    from oracle.schedulers import SD15_SCHEDULER_CONFIG, DPMSolverScheduler
    from oracle.unet import make_unet

    torch.set_num_threads(threads)
    net = make_unet(29)
    g = torch.Generator().manual_seed(29)
    pe, ne = torch.randn(1, 77, 768, generator=g), torch.randn(1, 77, 768, generator=g)
    lat = torch.randn(1, 4, 64, 64, generator=g)

    class Trunc(DPMSolverScheduler):            # run only the first n_steps of the 25-step schedule
        def set_timesteps(self, *a, **k):
            super().set_timesteps(*a, **k)
            self.timesteps = self.timesteps[:n_steps]

    sched = Trunc.from_config(SD15_SCHEDULER_CONFIG, solver_order=2, algorithm_type="dpmsolver++",
                              final_sigmas_type="zero")
    t0 = time.perf_counter()
    denoise(net, sched, pe, ne, lat, STEPS_PER_IMAGE, guidance_scale=GUIDANCE)
    return (time.perf_counter() - t0) / n_steps


def cpu_baseline(sample_steps=1):
    threads = len(os.sched_getaffinity(0))
    _oracle_cfg_step_seconds(1, threads)                       # warm-up (allocations, thread pool)
    s = _oracle_cfg_step_seconds(sample_steps, threads)
    return {"value": round(1.0 / (s * STEPS_PER_IMAGE), 6), "unit": "images/s", "cores": threads, "kind": "port",
            "sample": f"{sample_steps} CFG denoising step(s) at batch 1 (2 UNet sample-forwards each, fp32) of the "
                      f"25-step DPM-Solver++ schedule, {s:.2f} s/step, extrapolated x25",
            "gflops": round(2 * FLOP_PER_SAMPLE_FWD / s / 1e9, 1)}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    threads = len(os.sched_getaffinity(0))
    for _ in range(min(args.warmup, 1)):
        _oracle_cfg_step_seconds(1, threads)
    times = [_oracle_cfg_step_seconds(1, threads) for _ in range(max(1, min(args.steps, 3)))]
    s = statistics.mean(times)
    value = 1.0 / (s * STEPS_PER_IMAGE)
    line = {
        "impl": "reference", "metric": "images/sec (512x512, 25 steps DPM-Solver++, CFG 7.5)",
        "value": round(value, 6), "unit": "images/s", "n_gpus": int(os.environ.get("WORLD_SIZE", 1)),
        "steps": len(times), "warmup": min(args.warmup, 1), "ms_per_step": round(s * 1e3 * STEPS_PER_IMAGE, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(CONFIG),
        "cpu_baseline": {"value": round(value, 6), "unit": "images/s", "cores": threads, "kind": "port",
                         "sample": f"oracle restatement of the reference path (diffusers is not installable here), fp32, "
                                   f"all host threads: {len(times)} x 1 CFG denoising step at batch 1 (2 UNet "
                                   f"sample-forwards), mean {s:.2f} s/step, extrapolated x25 steps per image"},
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
