#!/usr/bin/env python
"""Two UNet steps + one VAE decode of the config-2 workload, launched kernel by kernel (no graph replay) so that
`ncu --metrics gpu__time_duration.sum` lists every launch of one step (see profiles/)."""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ldm_image_generator_b200 import DDPM, Decoder, UNet  # noqa: E402

B = int(os.environ.get("B", "64"))
torch.manual_seed(1234)
unet, dec = UNet().cuda().eval(), Decoder().cuda().eval()
ddpm = DDPM(model=unet)
x = torch.randn(B, 8, 32, 32, device="cuda")
unet._prepare(x.device).set_use_graphs(False)
random.seed(0)
z = ddpm.sample((B, 8, 32, 32), num_steps=int(os.environ.get("STEPS", "2")), x_T=x, progress=False)
img = dec.decode_to_uint8(z)
torch.cuda.synchronize()
print("ok", tuple(img.shape), unet._handle.launches, dec._handle.launches)
