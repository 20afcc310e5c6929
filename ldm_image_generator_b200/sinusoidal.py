"""Sinusoidal position / time encodings with the reference module API
(/root/reference/sinusoidal.py:6-41: ``PositionalEncoding2d``, ``TimeEncoding2d``).

The tables depend only on (C, H, W) or (C, t), so they are evaluated once on the host in
fp32 with the reference's own operation order (same torch CPU sin/cos, same bits) and handed
to the C library, which keeps them resident in HBM (`ldmb_unet_set_position_table`, `te_host`).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn


def position_table(channels: int, height: int, width: int) -> torch.Tensor:
    """[C, H, W] fp32 CPU.  sinusoidal.py:13-19: first C/2 channels vary along H, last C/2 along W;
    each half is [sin | cos] of coordinate/extent * pi * 2^-(k/(C/4))."""
    q = channels // 4
    freq = 1 / (2 ** (torch.arange(q).reshape(1, q, 1, 1) / q))
    rows = torch.arange(height, dtype=torch.float32).reshape(1, 1, height, 1) / height
    cols = torch.arange(width, dtype=torch.float32).reshape(1, 1, 1, width) / width
    enc_h = torch.cat([torch.sin(rows * math.pi * freq), torch.cos(rows * math.pi * freq)], dim=1)
    enc_w = torch.cat([torch.sin(cols * math.pi * freq), torch.cos(cols * math.pi * freq)], dim=1)
    full = torch.cat([enc_h.expand(1, 2 * q, height, width), enc_w.expand(1, 2 * q, height, width)], dim=1)
    return full[0].contiguous()


def time_table(channels: int, t: torch.Tensor, max_timesteps: int = 10000) -> torch.Tensor:
    """[len(t), C] fp32 CPU.  sinusoidal.py:32-39: raw integer t (not normalised) times pi times
    max_timesteps^-(k/(C/2)); first half sin, second half cos."""
    half = channels // 2
    t = t.detach().to("cpu")
    tt = t.reshape(-1, 1).expand(t.numel(), half)
    freq = (1 / (max_timesteps ** (torch.arange(half) / half))).unsqueeze(0)
    return torch.cat([torch.sin(tt * math.pi * freq), torch.cos(tt * math.pi * freq)], dim=1).float().contiguous()


class PositionalEncoding2d(nn.Module):
    def __init__(self, channels, return_encoding_only=False):
        super().__init__()
        self.channels = channels
        self.return_encoding_only = return_encoding_only

    def table(self, height: int, width: int) -> torch.Tensor:
        return position_table(self.channels, height, width)

    def forward(self, x):
        emb = self.table(x.shape[2], x.shape[3]).to(device=x.device, dtype=x.dtype).unsqueeze(0).expand(*x.shape)
        return emb if self.return_encoding_only else x + emb


class TimeEncoding2d(nn.Module):
    def __init__(self, channels, max_timesteps=10000, return_encoding_only=False):
        super().__init__()
        self.channels = channels
        self.max_timesteps = max_timesteps
        self.return_encoding_only = return_encoding_only

    def table(self, t: torch.Tensor) -> torch.Tensor:
        return time_table(self.channels, t, self.max_timesteps)

    # t: [batch_size]
    def forward(self, x, t):
        emb = self.table(t).to(device=x.device, dtype=x.dtype)[:, :, None, None].expand(*x.shape)
        return emb if self.return_encoding_only else x + emb
