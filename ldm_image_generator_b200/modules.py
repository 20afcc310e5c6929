"""Building blocks of the UNet with the reference module API (/root/reference/modules.py:7-36).

These classes are *parameter containers with the reference's names and shapes* (so
``state_dict`` / ``load_state_dict`` round-trip with reference checkpoints and a given torch
seed initialises the same weights).  Their arithmetic is not run module by module: the
whole UNet step executes inside libldmb200.so (``UNet.forward``).  Calling a sub-module on
its own raises, except ``ChannelNorm`` which maps to a single library kernel.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import runtime


class _FusedIntoUNet(nn.Module):
    def forward(self, *args, **kwargs):
        raise runtime.LdmbError(
            f"{type(self).__name__} is executed inside the fused UNet step of libldmb200 "
            "(call UNet.forward / DDPM.sample); it has no standalone or CPU implementation.")


class ReGLU(_FusedIntoUNet):
    """c(a(x) * relu(b(x))) with 1x1 convolutions (modules.py:8-15)."""

    def __init__(self, channels, ffn_mul=4):
        super().__init__()
        hidden = channels * ffn_mul
        self.a = nn.Conv2d(channels, hidden, 1, 1, 0)
        self.b = nn.Conv2d(channels, hidden, 1, 1, 0)
        self.act = nn.ReLU()
        self.c = nn.Conv2d(hidden, channels, 1, 1, 0)


class ChannelNorm(nn.Module):
    """Parameter-free per-pixel normalisation over channels, unbiased variance (modules.py:18-25)."""

    def __init__(self, channels, eps=1e-4):
        super().__init__()
        self.eps = eps
        self._handles = {}

    def forward(self, x):
        runtime._require_cuda(x, "ChannelNorm input")
        if abs(self.eps - 1e-4) > 1e-12:
            raise runtime.LdmbError("libldmb200 fixes eps=1e-4 (modules.py:19)")
        key = (x.device, runtime.default_precision())
        if key not in self._handles:
            self._handles[key] = runtime.Handle(x.device, key[1])
        h = self._handles[key]
        B, Cc, H, W = x.shape
        rows = runtime.f32c(x.permute(0, 2, 3, 1)).reshape(B * H * W, Cc)        # NHWC rows (layout only)
        film = torch.cat([torch.ones(1, Cc, device=x.device), torch.zeros(1, Cc, device=x.device)], dim=1)
        out = torch.empty(B * H * W, Cc, device=x.device,
                          dtype=torch.bfloat16 if h.precision == "bf16" else torch.float32)
        h.channelnorm_film(rows, film, out, B * H * W, Cc, 1)
        return out.reshape(B, H, W, Cc).permute(0, 3, 1, 2).to(x.dtype)


class RandomMoE(_FusedIntoUNet):
    """general(x) + e1(x) + e2(x), two of four experts drawn with Python's random per call (modules.py:28-36)."""

    def __init__(self, channels, ffn_mul=1, num_experts=4):
        super().__init__()
        if ffn_mul != 1 or num_experts != 4:
            raise runtime.LdmbError("libldmb200 implements RandomMoE as the reference instantiates it: ffn_mul=1, 4 experts")
        self.general = ReGLU(channels, ffn_mul=ffn_mul)
        self.experts = nn.ModuleList([ReGLU(channels, ffn_mul=ffn_mul) for _ in range(num_experts)])
