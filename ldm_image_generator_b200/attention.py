"""Attention modules with the reference API (/root/reference/attention.py:5-98).

Parameter containers only; window attention runs inside the fused UNet step
(csrc/kernels_simt.cu: window_attention_kernel + in/out projection GEMMs).
"""
from __future__ import annotations

import torch.nn as nn

from .modules import _FusedIntoUNet


class WindowAttention(_FusedIntoUNet):
    def __init__(self, channels=512, n_heads=8, window_size=4, shift=0):
        super().__init__()
        self.attention = nn.MultiheadAttention(channels, n_heads, batch_first=True)
        self.window_size = window_size
        self.shift = shift


class CrossAttention(_FusedIntoUNet):
    """Never executed by the reference either: UNet.forward drops ``condition`` (unet.py:93,101) and
    CrossAttention.forward returns None (attention.py:92-98).  Kept so checkpoints load strictly."""

    def __init__(self, channels=512, n_heads=8):
        super().__init__()
        self.attention = nn.MultiheadAttention(channels, n_heads, batch_first=True)
