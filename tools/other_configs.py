#!/usr/bin/env python
"""BASELINE.json configs[3] and configs[4] on one GPU (bounded): the wide UNet (2x base channels) at latent 64x64 and the
VAE encode+decode round trip at 512x512.  Prints timings; checks finiteness and the tcgen05 watchdog."""
import os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldm_image_generator_b200 import DDPM, Decoder, Encoder, UNet


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


torch.manual_seed(1234)
which = os.environ.get("WHICH", "wide,vae")
if "wide" in which:
    B, steps = int(os.environ.get("B", "16")), int(os.environ.get("STEPS", "20"))
    unet = UNet(channels=[256, 512, 1024, 2048]).cuda().eval()
    ddpm = DDPM(model=unet)
    x = torch.randn(B, 8, 64, 64, device="cuda")

    def run():
        random.seed(0)
        return ddpm.sample((B, 8, 64, 64), num_steps=steps, x_T=x, progress=False, schedule=list(range(0, 200, 200 // steps)))
    ms, z = timed(run, 2)
    assert torch.isfinite(z).all() and unet._handle.device_fault() == 0
    print(f"configs[3] wide UNet (1.53 B params), {B} images/GPU, latent 64x64: {ms / steps:8.2f} ms per UNet step "
          f"({214.82 * B / (ms / steps):7.1f} TFLOP/s on the hoisted algorithmic 214.82 GFLOP per image-step), latent std {float(z.std()):.3f}")
    del unet, ddpm
    torch.cuda.empty_cache()
if "full3" in which:
    # configs[3] as stated: 1000-step DDPM schedule (num_steps = num_timesteps), 16 images per GPU, then decode to 512x512
    B = int(os.environ.get("B", "16"))
    unet, dec = UNet(channels=[256, 512, 1024, 2048]).cuda().eval(), Decoder().cuda().eval()
    ddpm = DDPM(model=unet)
    x = torch.randn(B, 8, 64, 64, device="cuda")

    def run_full():
        z = ddpm.sample((B, 8, 64, 64), seed=0, num_steps=1000, x_T=x, progress=False)
        return dec.decode_to_uint8(z)
    ms, img = timed(run_full, 1)
    assert tuple(img.shape) == (B, 512, 512, 3) and unet._handle.device_fault() == 0 and dec._handle.device_fault() == 0
    assert 0 < float(img.float().std())
    print(f"configs[3] wide UNet, {B} images/GPU, 1000 DDIM steps + decode to 512x512: {ms / 1e3:7.2f} s per batch -> "
          f"{B / ms * 1e3:6.2f} images/s per GPU (x8 GPUs, no collective on the path: {8 * B / ms * 1e3:6.1f} images/s), "
          f"{(214.82 * 1000 + 322.34) * B / ms:7.1f} TFLOP/s")
    del unet, ddpm, dec
    torch.cuda.empty_cache()
if "vae" in which:
    Bv = int(os.environ.get("BV", "16"))
    enc, dec = Encoder().cuda().eval(), Decoder().cuda().eval()
    img = torch.randn(Bv, 3, 512, 512, device="cuda").clamp(-1, 1)
    with torch.no_grad():
        ms, out = timed(lambda: dec(enc(img)), 3)
    assert torch.isfinite(out).all() and enc._handle.device_fault() == 0 and dec._handle.device_fault() == 0
    gf = (312.59 + 322.34) * Bv
    print(f"configs[4] VAE encode+decode round trip at 512x512, micro-batch {Bv}: {ms:8.2f} ms -> {Bv / ms * 1e3:7.1f} images/s, "
          f"{gf / ms:7.1f} TFLOP/s; output {tuple(out.shape)}")
