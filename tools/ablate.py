#!/usr/bin/env python
"""In-graph cost of each kernel class of the UNet step: time the CUDA-graph replay of one step with the class's
launches dropped (ldmb_debug_skip_classes; results are garbage, timing is not) and subtract from the full step.
Unlike the ncu launch list (cold-cache, serialised) this sees warm L2 and programmatic-dependent-launch overlap."""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ldm_image_generator_b200 import DDPM, UNet  # noqa: E402

B = int(os.environ.get("B", "64"))
STEPS = int(os.environ.get("STEPS", "10"))
torch.manual_seed(1234)
unet = UNet().cuda().eval()
ddpm = DDPM(model=unet)
x = torch.randn(B, 8, 32, 32, device="cuda")
h = unet._prepare(x.device)


def run(skip):
    h.skip_classes(skip)
    ms = []
    for it in range(4):
        random.seed(0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ddpm.sample((B, 8, 32, 32), num_steps=STEPS, x_T=x, progress=False)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1) / STEPS)
    return min(ms[2:])      # iterations 0/1 launch eagerly and capture the graph


full = run([])
print(f"full UNet step (graph replay): {full * 1e3:9.1f} us")
classes = [c for c in h.PROFILE_CLASSES if c not in ("vae_conv3x3", "vae_gemm", "gemm_cuda_core")]
tot = 0.0
for c in classes:
    t = run([c])
    tot += full - t
    print(f"  without {c:20s} {t * 1e3:9.1f} us   -> class costs {1e3 * (full - t):8.1f} us ({100 * (full - t) / full:5.1f}%)")
print(f"  sum of class costs {tot * 1e3:9.1f} us; everything skipped: {run(classes) * 1e3:9.1f} us")
