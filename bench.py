#!/usr/bin/env python
"""Headline benchmark: images/sec of N-step LDM sampling + VAE decode (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch: ``DDPM.sample`` (50 DDIM iterations of the default
UNet at latent 8x32x32) followed by the VAE decode to 256x256 uint8 images, for 64 images per GPU
(BASELINE.json configs[1]; at N GPUs the batch is sharded by image, 64 per GPU, no collective on the
denoise path, so N=8 is configs[2]'s 512 images).  Random-init weights under torch.manual_seed(1234),
noise seed 0, eval mode, eta=0.

  value  images/sec with x_T already resident in HBM, CUDA events, max over ranks
  e2e    the same through the public module API with HOST buffers: pinned x_T -> device, sample, decode,
         uint8 images -> pinned host (+ the final NCCL all_gather of the images at N>1)
  roofline / kernels   per-kernel-class CUDA-event times of one extra (untimed) step, library-side events
  cpu_baseline         the CPU oracle port on this box's host cores, bounded sample, rank 0 at N=1 only

``--impl reference`` times the CPU implementation (oracle port of the reference's algorithm; the Python
reference itself cannot travel to the GPU box) on a bounded sample per step and prints the same line shape.
"""
from __future__ import annotations

import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "images/sec (50-step LDM sampling + VAE decode, 256x256)"
UNET_GFLOP_PER_IMAGE_STEP = 14.00      # hoisted algorithmic figure, SURVEY.md 8d
UNET_GFLOP_ENCODINGS_PER_STEP = 19.33  # batch-invariant Encodings MLP, once per step
DECODER_GFLOP_PER_IMAGE = 80.59


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="MEASURED_PEAKS.json (of measured)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="B200_PROFILING.md fallback (of fallback)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


# ----------------------------------------------------------------------------------------- CPU arm
def cpu_sample_once(sd_unet, ucfg, sd_dec, dcfg, R, num_steps: int, unet_batch: int, latent: int):
    """One bounded sample of the workload on the host: one UNet step at batch `unet_batch` and one decode of one
    image, extrapolated linearly (every step costs the same; nothing in the algorithm depends on batch size)."""
    x = torch.randn(unet_batch, ucfg.input_channels, latent, latent)
    t = torch.full((unet_batch,), 500, dtype=torch.long)
    plan = R.draw_plan(len(R.block_table(ucfg)), False)
    t0 = time.perf_counter()
    with torch.no_grad():
        R.unet_forward(sd_unet, ucfg, x, t, plan)
    t1 = time.perf_counter()
    with torch.no_grad():
        R.decoder_forward(sd_dec, dcfg, x[:1])
    t2 = time.perf_counter()
    per_image = num_steps * (t1 - t0) / unet_batch + (t2 - t1)
    return 1.0 / per_image, (t1 - t0), (t2 - t1)


def build_cpu_arm():
    from oracle import restate as R     # the CPU arm being timed / the checker -- never the product path
    from ldm_image_generator_b200 import Decoder, UNet
    torch.manual_seed(1234)
    u = UNet(); d = Decoder()
    sd_u = {k: v.detach() for k, v in u.state_dict().items()}
    sd_d = {k: v.detach() for k, v in d.state_dict().items()}
    return R, sd_u, R.UNetCfg(), sd_d, R.DecoderCfg()


def run_reference(args, rank: int):
    if rank != 0:
        return
    # all the host cores this process may use (torchrun exports OMP_NUM_THREADS=1 to its workers)
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:
        pass
    R, sd_u, ucfg, sd_d, dcfg = build_cpu_arm()
    random.seed(0); torch.manual_seed(0)
    ub = 4
    for _ in range(max(args.warmup, 0)):
        cpu_sample_once(sd_u, ucfg, sd_d, dcfg, R, args.num_steps, ub, args.latent)
    vals, t0 = [], time.perf_counter()
    for _ in range(args.steps):
        vals.append(cpu_sample_once(sd_u, ucfg, sd_d, dcfg, R, args.num_steps, ub, args.latent))
    wall = time.perf_counter() - t0
    v = sum(x[0] for x in vals) / len(vals)
    sample = (f"per step: 1 UNet forward at batch {ub} + 1 decode of 1 image, fp32, extrapolated linearly to "
              f"{args.num_steps} DDIM steps per image (UNet {vals[-1][1]:.3f} s, decode {vals[-1][2]:.3f} s)")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "images/sec", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": v, "unit": "images/sec", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    args._emit_ready()
    print(json.dumps(line), flush=True)


def workload_config(args, n):
    return {"workload": f"configs[1]: random-init LDM, default UNet (385.7M params) + default VAE Decoder (f=8), "
                        f"{args.batch} images/GPU at 256x256, {args.num_steps} DDIM steps, eta=0, {args.mode} mode",
            "global_batch": args.batch * n, "per_gpu_batch": args.batch, "latent": [8, args.latent, args.latent],
            "num_steps": args.num_steps, "sharding": "by image, no collective on the denoise path" if n > 1 else "single GPU",
            "l2": "no flush: 0.77 GB of bf16 weights + >100 MB of activations stream through the 126 MB L2 every UNet step"}


# ----------------------------------------------------------------------------------------- GPU arm
def run_ours(args, rank: int, world: int, local_rank: int):
    import torch.distributed as dist
    from ldm_image_generator_b200 import DDPM, Decoder, UNet, parallel
    assert torch.cuda.is_available(), "bench.py (impl ours) needs a CUDA device; there is no CPU fallback"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234)
    unet = UNet(); dec = Decoder()
    unet.to(dev); dec.to(dev)
    unet.train(args.mode == "train"); dec.eval()
    unet.set_precision(args.precision); dec.set_precision(args.precision)
    ddpm = DDPM(model=unet)
    B, L = args.batch, args.latent
    shape = (B, 8, L, L)
    # identical seeds on every rank; each rank takes its slice of the full noise batch
    x_host = parallel.shard_noise((B * world, 8, L, L), seed=0, rank=rank, world=world).pin_memory()
    x_dev = x_host.to(dev)
    out_host = torch.empty(B, 8 * L, 8 * L, 3, dtype=torch.uint8).pin_memory()

    def step_device():
        parallel.seed_plan_rng(0)       # same expert plan on every rank and every step (fixed workload)
        z = ddpm.sample(shape, seed=None, num_steps=args.num_steps, x_T=x_dev, progress=False)
        return dec.decode_to_uint8(z)

    def step_e2e():
        parallel.seed_plan_rng(0)
        xd = x_host.to(dev, non_blocking=True)
        z = ddpm.sample(shape, seed=None, num_steps=args.num_steps, x_T=xd, progress=False)
        u8 = dec.decode_to_uint8(z)
        if world > 1:
            parallel.gather_images(u8, B * world)      # the one collective of the path: final image gather over NVLink
        out_host.copy_(u8, non_blocking=True)
        torch.cuda.current_stream().synchronize()     # the caller holds the images when the step returns
        return out_host

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = unet._handle.launches + dec._handle.launches
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), unet._handle.launches + dec._handle.launches - l0

    clocks = ClockSampler(local_rank)
    timed(step_device, 0, max(args.warmup, 3))        # warm-up (also sizes workspaces, uploads weights)
    clocks.start()
    ms, launches = timed(step_device, args.steps, 0)
    clk = clocks.stop()
    ms_e2e, _ = timed(step_e2e, args.steps, 1)
    assert unet._handle.device_fault() == 0 and dec._handle.device_fault() == 0, "tcgen05 watchdog fired"

    # one extra, untimed step with library-side per-launch events -> per-kernel-class roofline
    pk = peaks()
    unet._handle.profile_begin(); dec._handle.profile_begin()
    step_device()
    pu, pd = unet._handle.profile_end(), dec._handle.profile_end()
    classes = {}
    for name in pu:
        m = pu[name][0] + pd[name][0]; w = pu[name][1] + pd[name][1]; n = pu[name][2] + pd[name][2]
        if n:
            classes[name] = {"ms": round(m, 3), "launches": n, "work": w}
    total_ms = sum(c["ms"] for c in classes.values()) or 1.0
    kernels = {}
    for name, c in classes.items():
        tensor = name in unet._handle.TENSOR_CLASSES
        ach = c["work"] / (c["ms"] * 1e-3) / (1e12 if tensor else 1e9) if c["ms"] > 0 and c["work"] > 0 else None
        kernels[name] = {"share": round(c["ms"] / total_ms, 4), "ms": c["ms"], "launches": c["launches"],
                         "achieved": None if ach is None else round(ach, 2), "unit": "TFLOP/s" if tensor else "GB/s"}
    dom = max((k for k in kernels if kernels[k]["achieved"] is not None), key=lambda k: kernels[k]["ms"])
    tensor = kernels[dom]["unit"] == "TFLOP/s"
    peak = pk["tf_sustained"] if tensor else pk["hbm"]
    # DRAM traffic of the dominant class's most frequent launch, from the committed `ncu --set full` capture
    traffic, traffic_note = None, None
    tp = os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")
    if os.path.exists(tp):
        t = json.load(open(tp)).get(dom)
        if t:
            traffic, traffic_note = t["dram_bytes_per_launch"], t["note"]
    roofline = {"kernel": dom, "bound": "tensor" if tensor else "hbm", "achieved": kernels[dom]["achieved"], "peak": peak,
                "unit": kernels[dom]["unit"], "frac": round(kernels[dom]["achieved"] / peak, 4), "traffic": traffic,
                "traffic_note": traffic_note,
                "peak_source": pk["source"] + (", sustained (kernel timed inside a long step)" if tensor else ""),
                "avg_launch_us": round(1000.0 * kernels[dom]["ms"] / kernels[dom]["launches"], 2)}

    # The per-launch events above serialise the launches (no programmatic-dependent-launch overlap) and include each
    # launch's latency.  What the dominant class costs INSIDE the graph-replayed step: the timed step with the class's
    # launches dropped (ldmb_debug_skip_classes: results are garbage, the launch sequence and timing are not), subtracted.
    if dom in unet._handle.PROFILE_CLASSES and not dom.startswith("vae"):
        t_full, _ = timed(step_device, 2, 1)
        unet._handle.skip_classes([dom])
        t_skip, _ = timed(step_device, 2, 2)
        unet._handle.skip_classes([])
        timed(step_device, 0, 1)
        marginal_ms = (t_full - t_skip) / 2
        if marginal_ms > 0:
            ach = classes[dom]["work"] / (marginal_ms * 1e-3) / (1e12 if tensor else 1e9)
            roofline["in_step"] = {"class_ms_per_step": round(marginal_ms, 3), "achieved": round(ach, 2),
                                   "frac": round(ach / peak, 4),
                                   "note": "step time minus the step time with this class's launches dropped (graph replay, "
                                           "PDL overlap and warm L2 as in the timed region)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    images = B * world * args.steps
    value = images / (ms * 1e-3)
    e2e_value = images / (ms_e2e * 1e-3)
    alg_tflop = (B * (args.num_steps * UNET_GFLOP_PER_IMAGE_STEP + DECODER_GFLOP_PER_IMAGE)
                 + args.num_steps * UNET_GFLOP_ENCODINGS_PER_STEP) / 1e3
    line = {"metric": METRIC, "value": round(value, 3), "unit": "images/sec", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
            "data": "synthetic", "config": workload_config(args, world),
            "e2e": {"value": round(e2e_value, 3), "unit": "images/sec", "h2d_bytes_per_step": x_host.numel() * 4,
                    "d2h_bytes_per_step": out_host.numel(), "ms_per_step": round(ms_e2e / args.steps, 3)},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "kernels": kernels,
            "whole_step": {"algorithmic_tflop_per_gpu_step": round(alg_tflop, 2),
                           "achieved_tflops_per_gpu": round(alg_tflop / (ms / args.steps * 1e-3), 1),
                           "frac_of_sustained_bf16_peak": round(alg_tflop / (ms / args.steps * 1e-3) / pk["tf_sustained"], 4)}}
    if world == 1 and not args.no_cpu_baseline:
        try:
            torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
        except Exception:
            pass
        R, sd_u, ucfg, sd_d, dcfg = build_cpu_arm()
        cpu_sample_once(sd_u, ucfg, sd_d, dcfg, R, args.num_steps, 4, L)
        t0, vals = time.perf_counter(), []
        while time.perf_counter() - t0 < 12.0 and len(vals) < 8:
            vals.append(cpu_sample_once(sd_u, ucfg, sd_d, dcfg, R, args.num_steps, 4, L))
        v = sum(x[0] for x in vals) / len(vals)
        line["cpu_baseline"] = {"value": round(v, 5), "unit": "images/sec", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{len(vals)} x (1 UNet forward at batch 4 + 1 decode of 1 image), fp32 oracle port with the "
                                          f"Encodings MLP evaluated once per batch, extrapolated to {args.num_steps} steps/image "
                                          f"(UNet {vals[-1][1]:.3f} s, decode {vals[-1][2]:.3f} s)"}
    args._emit_ready()
    print(json.dumps(line), flush=True)
    os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="eval", choices=["eval", "train"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=64, help="images per GPU")
    ap.add_argument("--num-steps", type=int, default=50)
    ap.add_argument("--latent", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    args._emit_ready = lambda: None
    if args.impl == "reference":
        run_reference(args, rank)
        return
    # keep stdout clean for the ONE JSON line: libraries (NCCL prints its version banner to stdout) go to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit_ready():
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
    args._emit_ready = emit_ready
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        emit_ready()
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
