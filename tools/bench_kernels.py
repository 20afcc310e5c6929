#!/usr/bin/env python
"""Per-shape timing of the tcgen05 GEMM / conv kernels through the C ABI (CUDA events, L2-cold-ish: rotating buffers).
Shapes are the ones the config-2 workload launches.  Prints one line per shape."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ldm_image_generator_b200 import runtime  # noqa: E402


def timeit(fn, iters=20):
    """Average GPU time per launch: `iters` launches captured in one CUDA graph (no host launch overhead)."""
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):
            fn(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for i in range(iters):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * iters) * 1e3   # us


def main():
    h = runtime.Handle(torch.device("cuda", 0), "bf16")
    B = int(os.environ.get("B", "64"))
    print(f"# batch {B}")
    gemms = []
    for lvl, C in enumerate((128, 256, 512, 1024)):
        M = B * (32 >> lvl) ** 2
        gemms += [("ffn_ab", M, 6 * C, C), ("ffn_c", M, C, 3 * C), ("qkv", M, 3 * C, C), ("c+out", M, C, 4 * C)]
    only = os.environ.get("ONLY")          # e.g. ONLY=ffn_ab:65536
    if only:
        nm, mm = only.split(":")
        gemms = [g for g in gemms if g[0] == nm and g[1] == int(mm)]
    NB = 4   # rotate over NB buffer sets so consecutive launches do not hit the same L2 lines
    for name, M, N, K in gemms:
        A = [torch.randn(M, K, device="cuda").bfloat16() for _ in range(NB)]
        W = [torch.randn(N, K, device="cuda").bfloat16() for _ in range(NB)]
        bias = torch.randn(N, device="cuda")
        out = [torch.empty(M, N, device="cuda", dtype=torch.bfloat16) for _ in range(NB)]
        us = timeit(lambda i: h.gemm(A[i % NB], W[i % NB], bias, out[i % NB], M, N, K))
        fl = 2.0 * M * N * K
        by = 2.0 * (M * K + N * K + M * N)
        lib = ""
        if os.environ.get("CUBLAS"):       # library reference point for the same shape (plain bf16 GEMM, no epilogue)
            us2 = timeit(lambda i: torch.matmul(A[i % NB], W[i % NB].t(), out=out[i % NB]))
            lib = f"   | cuBLAS {us2:7.1f} us {fl / us2 / 1e6:7.1f} TFLOP/s"
        print(f"gemm {name:8s} M={M:6d} N={N:5d} K={K:5d}  {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  {by / us / 1e3:7.1f} GB/s(alg){lib}")
    if only:
        return
    convs = [(B, 32, 32, 512, 512), (B, 64, 64, 256, 256), (B, 128, 128, 128, 128), (B, 256, 256, 64, 64)]
    for (b, H, W_, C, N) in convs:
        bb = max(1, min(b, 16))
        x = torch.randn(bb, H, W_, C, device="cuda").bfloat16()
        w = torch.randn(N, 9 * C, device="cuda").bfloat16()
        bias = torch.randn(N, device="cuda")
        out = torch.empty(bb, H, W_, N, device="cuda", dtype=torch.bfloat16)
        us = timeit(lambda i: h.conv3x3(x, w, bias, out, bb, H, W_, C, N, act=2))
        fl = 2.0 * bb * H * W_ * N * 9 * C
        print(f"conv3x3 B={bb} {H}x{W_} C={C} N={N}  {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s")
    assert h.device_fault() == 0


if __name__ == "__main__":
    main()
