"""Bulk image -> latent pre-encoding: the step before LDM training in the reference
(`LatentImageDataset.set_size`, /root/reference/dataset.py:119-175), with the VAE encoder run in
micro-batches on the GPU through ``ldmb_vae_encode`` instead of one image per call.

Wire format (unchanged, so `train_ldm.py`'s `LatentImageDataset.__getitem__`, dataset.py:177-186, reads it):
``<cache_dir>/<i>.pt`` = ``torch.save`` of the latent of image i as a ``[1, latent_channels, size/8, size/8]``
fp32 tensor on the encoder's device.

Host-side preprocessing restates dataset.py:136-166 (PIL, byte arithmetic):
  * the image is scaled so its longer side is ``size`` (integer truncation of the shorter side, NEAREST),
  * blurred with ``GaussianBlur(1)`` when the source is larger than half the scaled size in either direction,
  * pasted centred on a black ``size x size`` RGB canvas,
  * mapped to ``uint8 / 127.5 - 1`` fp32, CHW.
File discovery follows dataset.py:104-107: ``**/*.jpg`` recursively plus ``*.png`` in the top directory of every
source, then ``[:max_len]`` -- note that the reference's default ``max_len=-1`` therefore drops the LAST image.
"""
from __future__ import annotations

import glob
import os
import shutil
from typing import Iterable, List, Sequence

import numpy as np
import torch
from PIL import Image, ImageFile, ImageFilter

ImageFile.LOAD_TRUNCATED_IMAGES = True          # dataset.py:16-17


def list_images(source_dirs: Sequence[str], max_len: int | None = -1) -> List[str]:
    """dataset.py:104-107."""
    paths: List[str] = []
    for d in source_dirs:
        paths += glob.glob(os.path.join(d, "**/*.jpg"), recursive=True) + glob.glob(os.path.join(d, "*.png"))
    return paths[:max_len]


def preprocess_image(img: Image.Image, size: int) -> np.ndarray:
    """One image -> fp32 [3, size, size] in [-1, 1] (dataset.py:138-166)."""
    a, b = img.size                                # PIL: (width, height); the reference calls them H, W
    if a > b:
        b = int(b * size / a)
        a = size
    else:
        a = int(a * size / b)
        b = size
    blur = img.size[0] > a / 2 or img.size[1] > b / 2
    img = img.resize((a, b), Image.NEAREST)
    if blur:
        img = img.filter(ImageFilter.GaussianBlur(1))
    canvas = Image.new("RGB", (size, size), (0, 0, 0))
    canvas.paste(img, ((size - a) // 2, (size - b) // 2))
    arr = np.array(canvas.convert("RGB"))
    arr = arr / 127.5 - 1.0                        # float64 arithmetic, then cast: bit-identical to the reference
    return np.transpose(arr, (2, 0, 1)).astype(np.float32)


@torch.no_grad()
def encode_image_folder(source_dirs: Sequence[str], cache_dir: str, encoder: torch.nn.Module, size: int = 512,
                        max_len: int = -1, batch: int = 32, device: torch.device | str = "cuda",
                        paths: Iterable[str] | None = None) -> int:
    """Fill ``cache_dir`` with ``<i>.pt`` latents of every listed image; returns how many were written.
    ``encoder`` is this package's ``Encoder`` (or any module mapping ``[B,3,size,size] -> [B,C,size/8,size/8]``)."""
    paths = list(paths) if paths is not None else list_images(source_dirs, max_len)
    if os.path.exists(cache_dir):
        shutil.rmtree(cache_dir)                   # dataset.py:127-131: the cache is rebuilt from scratch
    os.mkdir(cache_dir)
    device = torch.device(device)
    encoder = encoder.to(device)
    staging = torch.empty(batch, 3, size, size, dtype=torch.float32)
    if device.type == "cuda":
        staging = staging.pin_memory()
    done = 0
    while done < len(paths):
        n = min(batch, len(paths) - done)
        for k in range(n):
            with Image.open(paths[done + k]) as im:
                staging[k].copy_(torch.from_numpy(preprocess_image(im, size)))
        z = encoder(staging[:n].to(device, non_blocking=True))
        torch.cuda.synchronize(device) if device.type == "cuda" else None     # the staging buffer is free again afterwards
        handle = getattr(encoder, "_handle", None)
        if handle is not None:
            handle.raise_on_fault()        # never write latents of a faulted kernel into the cache
        for k in range(n):
            torch.save(z[k:k + 1].clone(), os.path.join(cache_dir, f"{done + k}.pt"))
        done += n
    return done


class LatentCache(torch.utils.data.Dataset):
    """Reader of the cache (dataset.py:177-189): item i is the ``[C, h, w]`` latent of image i."""

    def __init__(self, cache_dir: str):
        self.cache_dir = cache_dir

    def __getitem__(self, index):
        try:
            z = torch.load(os.path.join(self.cache_dir, f"{index}.pt"))
        except Exception:
            print("Skipped error")
            z = torch.load(os.path.join(self.cache_dir, "0.pt"))
        return z[0]

    def __len__(self):
        return len(os.listdir(self.cache_dir))
