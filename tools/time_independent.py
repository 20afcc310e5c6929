import sys, time; sys.path.insert(0, "/root/repo")
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
