#!/usr/bin/env python
"""Device timeline of the CUDA-graph replay of UNet steps through CUPTI (torch.profiler): per-kernel duration inside
the graph (warm L2, programmatic dependent launch) and the idle gap before each kernel.  Under the profiler the
numbers are for attribution only (never a bench value)."""
import collections
import os
import random
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from ldm_image_generator_b200 import DDPM, UNet  # noqa: E402

B = int(os.environ.get("B", "64"))
STEPS = int(os.environ.get("STEPS", "6"))
torch.manual_seed(1234)
unet = UNet().cuda().eval()
ddpm = DDPM(model=unet)
x = torch.randn(B, 8, 32, 32, device="cuda")
for _ in range(2):
    random.seed(0)
    ddpm.sample((B, 8, 32, 32), num_steps=STEPS, x_T=x, progress=False)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    random.seed(0)
    ddpm.sample((B, 8, 32, 32), num_steps=STEPS, x_T=x, progress=False)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "emcpy" not in e.name and "emset" not in e.name]
ev.sort(key=lambda e: e.time_range.start)


def short(n):
    n = n.replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    n = re.sub(r"\(.*", "", n)
    return n.replace("__nv_bfloat16", "bf16")[:44]


# one step = from a stem kernel (stem_mma_kernel; pointwise_in_kernel<float> in fp32 mode) ... take the last full step
first = [i for i, e in enumerate(ev) if "stem_mma_kernel" in e.name or "pointwise_in_kernel<float" in e.name]      # the stem opens every step
lo, hi = first[-2], first[-1]
step = ev[lo:hi]
t0 = step[0].time_range.start
span = ev[hi].time_range.start - t0
busy = sum(e.time_range.end - e.time_range.start for e in step)
print(f"one step: {len(step)} kernels, span {span:.1f} us, sum of kernel durations {busy:.1f} us, idle {span - busy:.1f} us")
agg = collections.OrderedDict()
prev_end = None
for e in step:
    d = e.time_range.end - e.time_range.start
    gap = 0.0 if prev_end is None else max(0.0, e.time_range.start - prev_end)
    prev_end = max(prev_end or 0, e.time_range.end)
    a = agg.setdefault(short(e.name), [0, 0.0, 0.0])
    a[0] += 1; a[1] += d; a[2] += gap
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:46s} {a[0]:4d} x  {a[1]:8.1f} us  avg {a[1] / a[0]:6.1f}  gap-before avg {a[2] / a[0]:5.2f} us")
if os.environ.get("LIST"):
    for e in step:
        print(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:7.1f}  {short(e.name)}")

