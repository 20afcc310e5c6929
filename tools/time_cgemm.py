#!/usr/bin/env python
"""Graph-timed residual (accumulate-into-x, split-K) GEMMs at the deep-level shapes of config 2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ldm_image_generator_b200 import runtime
from tools.bench_kernels import timeit

h = runtime.Handle(torch.device("cuda", 0), "bf16")
for (M, N, K) in ((4096, 512, 1536), (4096, 512, 2048), (1024, 1024, 3072), (1024, 1024, 4096)):
    A = [torch.randn(M, K, device="cuda").bfloat16() for _ in range(3)]
    W = torch.randn(N, K, device="cuda").bfloat16()
    b = torch.zeros(N, device="cuda")
    x = [torch.randn(M, N, device="cuda") for _ in range(3)]
    us = timeit(lambda i: h.gemm(A[i % 3], W, b, x[i % 3], M, N, K, out_f32=2))
    us2 = timeit(lambda i: torch.matmul(A[i % 3], W.t()))
    print(f"x += A.W^T  M={M} N={N} K={K}: {us:6.1f} us ({2.0 * M * N * K / us / 1e6:6.1f} TFLOP/s)   cuBLAS bf16-out GEMM {us2:6.1f} us")
assert h.device_fault() == 0
