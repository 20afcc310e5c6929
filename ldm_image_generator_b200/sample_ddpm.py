#!/usr/bin/env python
"""Pixel-space DDPM sampling like /root/reference/sample_ddpm.py.  The reference script crashes at its own defaults
(DDPM() builds an 8-channel UNet but asks for 3-channel images, sample_ddpm.py:19,36); here the UNet is built with
input_channels=3 as config 1 of the benchmark does."""
import os
import sys

import numpy as np
import torch
from PIL import Image

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldm_image_generator_b200 import DDPM, UNet  # noqa: E402

ddpm_path = "./ddpm.pt"
image_size = 32
result_dir = "./ddpm_outputs/"
num_images = 10


def main():
    if not torch.cuda.is_available():
        print("Error: cuda is not available in this environment.")
        sys.exit(1)
    ddpm = DDPM(model=UNet(input_channels=3))
    if os.path.exists(ddpm_path):
        ddpm.load_state_dict(torch.load(ddpm_path))
        print("DDPM Model Loaded.")
    ddpm.to(torch.device('cuda'))
    os.makedirs(result_dir, exist_ok=True)
    for i in range(num_images):
        img = ddpm.sample((1, 3, image_size, image_size), seed=i)
        img = torch.clamp(img, -1, 1)
        arr = (img[0].cpu().numpy() * 127.5 + 127.5).astype(np.uint8).transpose(1, 2, 0)
        Image.fromarray(arr, mode='RGB').save(os.path.join(result_dir, f"{i}.jpg"))


if __name__ == "__main__":
    main()
