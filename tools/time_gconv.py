#!/usr/bin/env python
"""Graph-timed grouped 3x3 halo conv at the four level shapes of config 2 (LDMB_GCONV_DBG bits for A/B runs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldm_image_generator_b200 import runtime
from tools.bench_kernels import timeit
h = runtime.Handle(torch.device("cuda", 0), "bf16")
for lvl in range(4):
    B, C, H = 64, 128 << lvl, 32 >> lvl
    xm = [torch.randn(B, H, H, C, device="cuda").bfloat16() for _ in range(3)]
    x = [torch.randn(B, H, H, C, device="cuda") for _ in range(3)]
    w = torch.randn(C, 576, device="cuda").bfloat16(); b = torch.zeros(C, device="cuda")
    us = timeit(lambda i: h.grouped_conv3x3(xm[i % 3], w, b, x[i % 3], B, H, H, C))
    ug = timeit(lambda i: h.grouped_conv3x3(xm[i % 3], w, b, x[i % 3], B, H, H, C, force_generic=True))
    print(f"gconv level {lvl} (C={C}, {H}x{H}): halo-patch kernel {us:6.1f} us | generic 9-tap-load implicit GEMM {ug:6.1f} us")
assert h.device_fault() == 0
