"""VAE encoder / decoder with the reference API (/root/reference/vae.py:28-132), executed by libldmb200.so.

``Decoder.forward(z)`` and ``Encoder.forward(x)`` are one C-ABI call each (``ldmb_vae_decode`` /
``ldmb_vae_encode``): dense 3x3 convolutions as TMA-fed tcgen05 implicit GEMMs with fused
bias + LeakyReLU (+ skip) epilogues, ConvTranspose 2x2 as a GEMM with a scatter epilogue, and the
progressive-RGB head (to_rgb + bilinear x2 + add) as one HBM-bound kernel per level.
The training-only parts of vae.py (VectorQuantizer, Discriminator, losses) are out of scope.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import _lib, runtime
from .modules import _FusedIntoUNet


class ResBlock(_FusedIntoUNet):
    def __init__(self, channels):
        super().__init__()
        self.c1 = nn.Conv2d(channels, channels, 3, 1, 1)
        self.c2 = nn.Conv2d(channels, channels, 3, 1, 1)


class ResStack(_FusedIntoUNet):
    def __init__(self, channels, num_layers=2):
        super().__init__()
        self.seq = nn.Sequential(*[ResBlock(channels) for _ in range(num_layers)])


class DecoderStack(_FusedIntoUNet):
    def __init__(self, channels, num_layers, output_channels=3):
        super().__init__()
        self.layers = nn.Sequential(*[ResBlock(channels) for _ in range(num_layers)])
        self.to_rgb = nn.Conv2d(channels, output_channels, 1, 1, 0)


class _VaeNet(nn.Module):
    _which = -1

    def _init_runtime(self):
        self.precision = runtime.default_precision()
        self._handle: Optional[runtime.Handle] = None

    def set_precision(self, precision: str):
        if precision not in runtime.PRECISIONS:
            raise ValueError(precision)
        if precision != self.precision:
            self.precision, self._handle = precision, None
        return self

    def __getstate__(self):
        """copy.deepcopy / pickle / torch.save(module): the device handle (ctypes pointers) is per-object state that is
        rebuilt on first use -- the reference modules support all three."""
        state = self.__dict__.copy()
        state["_handle"] = None
        return state

    def invalidate_weights(self) -> None:
        """Force a re-upload of every parameter at the next call.  Needed after writes THROUGH ``.data``
        (``p.data.copy_()``, EMA code), which do not bump ``p._version`` and so are not detected."""
        if self._handle is not None:
            self._handle._param_keys.clear()

    def _prepare(self, device: torch.device) -> runtime.Handle:
        h = self._handle
        if h is None or h.device != torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device()):
            h = runtime.Handle(device, self.precision)
            h.vae_configure(self._which, self.image_channels, self.latent_channels, self.level_channels, self.level_blocks)
            self._handle = h
        h.vae_load(self._which, self.state_dict(keep_vars=True).items())
        return h


class Encoder(_VaeNet):
    _which = _lib.VAE_ENCODER

    def __init__(self, input_channels=3, latent_channels=8, channels=[64, 128, 256, 512], stages=[2, 2, 2, 2]):
        super().__init__()
        self.image_channels, self.latent_channels = input_channels, latent_channels
        self.level_channels, self.level_blocks = list(channels), list(stages)
        self.input_layer = nn.Conv2d(input_channels, channels[0], 1, 1, 0)
        self.output_layer = nn.Conv2d(channels[-1], latent_channels, 1, 1, 0)
        self.stages = nn.ModuleList([ResStack(c, l) for c, l in zip(channels, stages)])
        self.downsamples = nn.ModuleList([])
        for i, c in enumerate(channels):
            if i == len(self.stages) - 1:
                self.downsamples.append(nn.Identity())
            else:
                self.downsamples.append(nn.Sequential(nn.AvgPool2d(kernel_size=2), nn.Conv2d(c, channels[i + 1], 1, 1, 0)))
        self._init_runtime()

    def forward(self, x):
        runtime._require_cuda(x, "Encoder input")
        x = runtime.f32c(x)
        if x.dim() != 4 or x.shape[1] != self.image_channels:
            raise RuntimeError(f"expected input with {self.image_channels} channels, got shape {tuple(x.shape)}")
        f = 1 << (len(self.level_channels) - 1)
        with torch.cuda.device(x.device):
            h = self._prepare(x.device)
            z = torch.empty(x.shape[0], self.latent_channels, x.shape[2] // f, x.shape[3] // f, device=x.device)
            h.vae_encode(x, z)
            h.raise_on_fault()
        return z


class Decoder(_VaeNet):
    _which = _lib.VAE_DECODER

    def __init__(self, output_channels=3, latent_channels=8, channels=[512, 256, 128, 64], stages=[2, 2, 2, 2]):
        super().__init__()
        self.image_channels, self.latent_channels = output_channels, latent_channels
        self.level_channels, self.level_blocks = list(channels), list(stages)
        self.input_layer = nn.Conv2d(latent_channels, channels[0], 1, 1, 0)
        self.output_layer = nn.Conv2d(channels[-1], output_channels, 1, 1, 0)   # in the state_dict, never applied (vae.py:113)
        self.stages = nn.ModuleList([DecoderStack(c, l, output_channels=output_channels) for c, l in zip(channels, stages)])
        self.upsamples = nn.ModuleList([])
        for i, c in enumerate(channels):
            if i == 0:
                self.upsamples.append(nn.Identity())
            else:
                self.upsamples.append(nn.ConvTranspose2d(channels[i - 1], c, 2, 2, 0))
        self._init_runtime()

    def _decode(self, z, want_f32: bool, want_u8: bool):
        runtime._require_cuda(z, "Decoder input")
        z = runtime.f32c(z)
        if z.dim() != 4 or z.shape[1] != self.latent_channels:
            raise RuntimeError(f"expected latent with {self.latent_channels} channels, got shape {tuple(z.shape)}")
        f = 1 << (len(self.level_channels) - 1)
        B, _, hl, wl = z.shape
        with torch.cuda.device(z.device):
            h = self._prepare(z.device)
            img = torch.empty(B, self.image_channels, hl * f, wl * f, device=z.device) if want_f32 else None
            u8 = torch.empty(B, hl * f, wl * f, self.image_channels, device=z.device, dtype=torch.uint8) if want_u8 else None
            h.vae_decode(z, img, u8)
            h.raise_on_fault()
        return img, u8

    def forward(self, x):
        return self._decode(x, True, False)[0]

    def decode_to_uint8(self, z):
        """Decoder + the script's post-processing (sample_ldm.py:75-77) in one call: returns the
        [B, H, W, 3] uint8 image (clamp, *127.5+127.5, truncate, CHW->HWC) without the fp32 round trip."""
        return self._decode(z, False, True)[1]


class VAE(nn.Module):
    def __init__(self, encoder, decoder, quantizer=None):
        super().__init__()
        self.encoder = encoder
        self.decoder = decoder
        self.quantizer = quantizer

    def calclate_loss(self, x, noise_gain=0.1):
        raise NotImplementedError("training losses (vae.py:37-43) are outside the sampling path this package implements")

    @torch.no_grad()
    def encode(self, x):
        return self.encoder(x)      # no quantisation on this path (vae.py:45-48)

    @torch.no_grad()
    def decode(self, z):
        return self.decoder(z)
