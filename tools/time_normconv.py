#!/usr/bin/env python
"""Fused norm + FiLM + grouped conv kernel vs the separate norm and conv kernels at the deep-level shapes (graph replay)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from ldm_image_generator_b200 import runtime
h = runtime.Handle(torch.device("cuda", 0), "bf16")

def pack(w):
    C = w.shape[0]
    out = torch.zeros(C // 64, 64, 9, 64, device=w.device, dtype=w.dtype)
    wp = w.reshape(C // 64, 2, 32, 32, 9)
    for gl in range(2):
        out[:, gl * 32:(gl + 1) * 32, :, gl * 32:(gl + 1) * 32] = wp[:, gl].permute(0, 1, 3, 2)
    return out.reshape(C, 9 * 64).contiguous()

def timeit(f, n=20):
    for _ in range(3): f()
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        f()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n): f()
    torch.cuda.synchronize(); g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

for (B, H, W, C) in [(64, 8, 8, 512), (64, 4, 4, 1024), (8, 8, 8, 512), (64, 8, 8, 256)]:
    x = torch.randn(B, H, W, C, device="cuda"); film = torch.randn(H * W, 2 * C, device="cuda")
    w = pack((torch.randn(C, 32, 3, 3, device="cuda") / 17).bfloat16()); b = torch.randn(C, device="cuda")
    xm = torch.empty(B, H, W, C, device="cuda", dtype=torch.bfloat16)
    t_f = timeit(lambda: h.normconv(x, film, xm, w, b, B, H, W, C))
    t_n = timeit(lambda: h.channelnorm_film(x.reshape(-1, C), film, xm.reshape(-1, C), B * H * W, C, H * W))
    t_c = timeit(lambda: h.grouped_conv3x3(xm, w, b, x, B, H, W, C))
    print(f"B={B} {H}x{W} C={C}: fused {t_f:7.1f} us | norm {t_n:6.1f} + conv {t_c:6.1f} us", flush=True)
    assert h.device_fault() == 0
