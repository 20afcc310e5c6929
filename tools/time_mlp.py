#!/usr/bin/env python
"""Graph-timed fused feed-forward kernel at the config-2 level shapes (events around a CUDA graph of 20 launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldm_image_generator_b200 import runtime
from tools.bench_kernels import timeit

h = runtime.Handle(torch.device("cuda", 0), "bf16")
for lvl in (0, 1, 2):
    Cc, M = 128 << lvl, 65536 >> (2 * lvl)
    xm = [torch.randn(M, Cc, device="cuda").bfloat16() for _ in range(3)]
    x = [torch.randn(M, Cc, device="cuda") for _ in range(3)]
    w_ab = torch.randn(10 * Cc, Cc, device="cuda").bfloat16(); w_c = torch.randn(5 * Cc, Cc, device="cuda").bfloat16()
    b_ab = torch.zeros(10 * Cc, device="cuda"); b_c = torch.zeros(5 * Cc, device="cuda")
    try:
        us = timeit(lambda i: h.mlp_fused(xm[i % 3], w_ab, b_ab, w_c, b_c, x[i % 3], M, Cc, 1, 2))
        print(f"mlp_fused C={Cc} M={M}: {us:7.1f} us   {2.0 * M * 9 * Cc * Cc / us / 1e6:7.1f} TFLOP/s")
    except Exception as e:
        print("C", Cc, "failed:", str(e)[:100])
assert h.device_fault() == 0
