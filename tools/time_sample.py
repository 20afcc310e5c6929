#!/usr/bin/env python
"""Events around DDPM.sample((B, 8, 32, 32), 50 steps) -- the UNet part of the config-2 batch (B=64; env B overrides) -- for quick A/B of env knobs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldm_image_generator_b200 import DDPM, UNet
torch.manual_seed(1234)
u = UNet().cuda().eval(); d = DDPM(model=u)
B = int(os.environ.get("B", "64"))
x = torch.randn(B, 8, 32, 32, device="cuda")
fn = lambda: d.sample((B, 8, 32, 32), seed=0, num_steps=50, x_T=x, progress=False)
for _ in range(3): fn()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(f"sample {B} x 50 steps: median {sorted(ts)[2]:.2f} ms  min {min(ts):.2f}  ({sorted(ts)[2] / 50:.4f} ms per step)")
assert u._handle.device_fault() == 0
