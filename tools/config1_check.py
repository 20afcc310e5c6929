#!/usr/bin/env python
"""BASELINE.json configs[0] (the reference's own CPU-runnable case, as it has to be run: DDPM(UNet(input_channels=3)),
3x32x32, batch 4, 50 steps): the GPU sampler against the CPU oracle over the WHOLE trajectory, fp32 validation mode and
bf16, teacher-forced per-step errors included.  One-off check (the oracle needs ~1 minute on 16 host threads)."""
import os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import restate as R            # checker only
from ldm_image_generator_b200 import DDPM, UNet

cfg = R.UNetCfg(input_channels=3)
sd = R.make_unet_state(cfg, 1234)
torch.manual_seed(0)
x_T = torch.randn(4, 3, 32, 32)
steps = [int(v) for v in torch.linspace(0, 999, 50).int().numpy()]
t0 = time.perf_counter()
traj = []
want = R.ddim_sample(sd, cfg, x_T, steps, False, py_seed=0, trajectory=traj)
cpu_s = time.perf_counter() - t0
print(f"CPU oracle: {cpu_s:.1f} s for 4 images x 50 steps on {torch.get_num_threads()} threads -> {4 / cpu_s:.4f} images/s")
for prec in ("fp32", "bf16"):
    unet = UNet(input_channels=3).cuda().eval()
    unet.load_state_dict(sd); unet.set_precision(prec)
    ddpm = DDPM(model=unet)
    ddpm.sample((4, 3, 32, 32), seed=0, num_steps=50, x_T=x_T, progress=False)        # warm-up / graph capture
    torch.cuda.synchronize(); t0 = time.perf_counter()
    got = ddpm.sample((4, 3, 32, 32), seed=0, num_steps=50, x_T=x_T, progress=False).cpu()
    gpu_s = time.perf_counter() - t0
    # teacher-forced: every step's eps on the oracle trajectory's x
    worst = 0.0
    n_blocks = len(R.block_table(cfg))
    random.seed(0)
    for (t, x_in, eps) in traj:
        plan = R.draw_plan(n_blocks, False)
        e = unet._run(x_in.cuda(), [t] * 4, plan=plan).cpu()
        worst = max(worst, R.rel_l2(e, eps))
    print(f"{prec}: final x0 rel-L2 {R.rel_l2(got, want):.3e} (std {float(want.std()):.3g}); worst teacher-forced step rel-L2 {worst:.3e}; "
          f"GPU {gpu_s * 1e3:.1f} ms -> {4 / gpu_s:.1f} images/s")
    assert unet._handle.device_fault() == 0
