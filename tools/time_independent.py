#!/usr/bin/env python
"""Throughput of DDPM.sample_independent (per-image plans: the reference scripts' batch-1 loops computed as one batch) against the
shared-plan batch and the batch-1 loop through this library (profiles/r1_per_image_plans_throughput.txt)."""
import os, sys, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, random
from ldm_image_generator_b200 import DDPM, UNet
torch.manual_seed(1234)
u = UNet().cuda().eval(); d = DDPM(model=u)
for name, fn in (("shared plan  sample((64,...))", lambda: d.sample((64, 8, 32, 32), num_steps=50, progress=False)),
                 ("per-image    sample_independent(64)", lambda: d.sample_independent(64, (1, 8, 32, 32), num_steps=50))):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(2): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 2
    print(f"{name}: {dt * 1e3:7.1f} ms for 64 latents x 50 steps -> {64 / dt:6.1f} latents/s")
d2 = DDPM(model=u)
t0 = time.perf_counter(); d2.sample((1, 8, 32, 32), num_steps=50, progress=False); d2.sample((1, 8, 32, 32), num_steps=50, progress=False)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(4): d2.sample((1, 8, 32, 32), num_steps=50, progress=False)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 4
print(f"batch-1 loop (the reference scripts' structure, this library): {dt * 1e3:7.1f} ms per latent -> {1 / dt:6.1f} latents/s")
# device time of one per-image step against one shared-plan step (same batch, same t), per profile class
import numpy as np
x = torch.randn(64, 8, 32, 32, device="cuda")
u.precompute_film(x, [500])
co, _ = d.ddim_scalars(torch.cumprod(1 - d.beta, dim=0), 500, 480, 0)
plans = u.draw_plans(64)
shared = u.draw_plan()
for name, kw in (("shared-plan step", dict(plan=shared)), ("per-image step", dict(plans_per_image=plans))):
    for _ in range(3): u._run(x, [500] * 64, coef=co, out=x.clone(), check_params=False, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): u._run(x, [500] * 64, coef=co, out=x.clone(), check_params=False, **kw)
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 20:6.3f} ms on the device")
t0 = time.perf_counter(); u.draw_plans(64 * 50); print(f"drawing 64 x 50 plans: {(time.perf_counter() - t0) * 1e3:.1f} ms of host time")
