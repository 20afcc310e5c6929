#!/usr/bin/env python
"""LDM sampling CLI with the flags of the reference script (/root/reference/sample_ldm.py:11-23), running the
UNet / DDIM / VAE-decode path in libldmb200.  Extras: ``--batch`` denoises several images together -- with per-image
stochastic-depth / expert decisions and per-image noise drawn in the order the reference's batch-1 loop draws them
(DDPM.sample_independent), so the images are the ones the loop would produce; ``--shared-plan`` uses one decision per
step for the whole batch instead (faster: one reference forward on a batch); ``--precision`` picks bf16 (default) or
the fp32 validation mode."""
import argparse
import os
import sys

import numpy as np
import torch
from PIL import Image

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldm_image_generator_b200 import DDPM, Decoder  # noqa: E402


def main():
    parser = argparse.ArgumentParser(description="Sample LDM")
    parser.add_argument('-dp', '--ddpmpath', default='./ddpm.pt')
    parser.add_argument('-decp', '--decpath', default='./vae_decoder.pt')
    parser.add_argument('-d', '--device', default='cuda', choices=['cuda'],
                        help="this implementation is CUDA (sm_100a) only; there is no cpu/mps path")
    parser.add_argument('-fp16', default=False, type=bool, help="accepted for compatibility; see --precision")
    parser.add_argument('-s', '--size', default=512, type=int)
    parser.add_argument('-n', '--numimages', default=1, type=int)
    parser.add_argument('-t', '--timesteps', default=20, type=int)
    parser.add_argument('--seed', default=0, type=int)
    parser.add_argument('--batch', default=1, type=int, help="images denoised together (reference: 1)")
    parser.add_argument('--shared-plan', action='store_true', help="one stochastic-depth / expert decision per step for the whole batch")
    parser.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    parser.add_argument('-o', '--outdir', default='./ddpm_outputs/')
    args = parser.parse_args()

    if not torch.cuda.is_available():
        print("Error: cuda is not available in this environment.")
        sys.exit(1)
    device = torch.device(args.device)
    ddpm, decoder = DDPM(), Decoder()
    if os.path.exists(args.ddpmpath):
        ddpm.load_state_dict(torch.load(args.ddpmpath, map_location='cpu'))
        print("DDPM Model Loaded.")
    if os.path.exists(args.decpath):
        decoder.load_state_dict(torch.load(args.decpath, map_location='cpu'))
        print("VAE Decoder Loaded.")
    ddpm, decoder = ddpm.to(device), decoder.to(device)      # no .eval(): like the reference, stochastic depth stays live
    ddpm.model.set_precision(args.precision); decoder.set_precision(args.precision)
    os.makedirs(args.outdir, exist_ok=True)

    latent = args.size // 8                                  # latent_space_downscale_ratio (sample_ldm.py:28,66)
    torch.manual_seed(args.seed)
    torch.cuda.manual_seed(args.seed)
    done = 0
    while done < args.numimages:
        b = min(args.batch, args.numimages - done)
        if b == 1 or args.shared_plan:
            z = ddpm.sample((b, 8, latent, latent), seed=None, num_steps=args.timesteps, use_autocast=args.fp16)
        else:       # == [ddpm.sample((1, 8, latent, latent), ...) for _ in range(b)] of sample_ldm.py:71-72, as one batch
            z = ddpm.sample_independent(b, (1, 8, latent, latent), seeds=None, num_steps=args.timesteps, use_autocast=args.fp16)
        imgs = decoder.decode_to_uint8(z).cpu().numpy()       # clamp, *127.5+127.5, uint8, HWC fused into the last kernel
        for k in range(b):
            Image.fromarray(imgs[k], mode='RGB').save(os.path.join(args.outdir, f"{done + k}.jpg"))
        done += b


if __name__ == "__main__":
    main()
