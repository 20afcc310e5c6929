#!/usr/bin/env python
"""Where does the time go inside one tcgen05 GEMM launch?  Per-CTA %globaltimer stamps (debug hook)."""
import ctypes as C
import ctypes as C_
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from ldm_image_generator_b200 import runtime  # noqa: E402

NAMES = ["entry", "setup", "prev_done", "tma0", "data0", "mma_end", "acc0", "accN", "epi_end", "exit"]


def gconv(h, lvl):
    B, C, H = 64, 128 << lvl, 32 >> lvl
    xm = torch.randn(B, H, H, C, device="cuda").bfloat16(); x = torch.randn(B, H, H, C, device="cuda")
    w = torch.randn(C, 576, device="cuda").bfloat16(); b = torch.zeros(C, device="cuda")
    for _ in range(3):
        h.grouped_conv3x3(xm, w, b, x, B, H, H, C)
    torch.cuda.synchronize()
    h.lib.ldmb_debug_tc_trace(h.h, 1, None, 0)
    for _ in range(4):
        h.grouped_conv3x3(xm, w, b, x, B, H, H, C)
    buf = (C_.c_int64 * (16 * 256))()
    h.lib.ldmb_debug_tc_trace(h.h, 1, buf, 256)
    h.lib.ldmb_debug_tc_trace(h.h, 0, None, 0)
    a = np.frombuffer(buf, dtype=np.int64).reshape(256, 16)[:148, :11].astype(np.float64)
    a = a[a[:, 0] > 0]
    rel = (a - a[:, 0].min()) / 1000.0
    print(f"gconv level {lvl}: {len(a)} CTAs, kernel span {rel[:, 9].max():.2f} us")
    for i, nm in enumerate(NAMES + ["weights"]):
        col = rel[:, i]
        print(f"   {nm:10s} min {col.min():7.2f}  median {np.median(col):7.2f}  max {col.max():7.2f} us")


def mlp(h, lvl):
    Cc, M = 128 << lvl, 65536 >> (2 * lvl)
    xm = torch.randn(M, Cc, device="cuda").bfloat16(); x = torch.randn(M, Cc, device="cuda")
    w_ab = torch.randn(10 * Cc, Cc, device="cuda").bfloat16(); w_c = torch.randn(5 * Cc, Cc, device="cuda").bfloat16()
    b_ab = torch.zeros(10 * Cc, device="cuda"); b_c = torch.zeros(5 * Cc, device="cuda")
    for _ in range(3):
        h.mlp_fused(xm, w_ab, b_ab, w_c, b_c, x, M, Cc, 1, 2)
    torch.cuda.synchronize()
    h.lib.ldmb_debug_tc_trace(h.h, 1, None, 0)
    for _ in range(4):
        h.mlp_fused(xm, w_ab, b_ab, w_c, b_c, x, M, Cc, 1, 2)
    buf = (C_.c_int64 * (16 * 256))()
    h.lib.ldmb_debug_tc_trace(h.h, 1, buf, 256)
    h.lib.ldmb_debug_tc_trace(h.h, 0, None, 0)
    raw = np.frombuffer(buf, dtype=np.int64).reshape(256, 16)[:148].astype(np.float64)
    raw = raw[raw[:, 0] > 0]
    rel = (raw[:, :10] - raw[:, 0].min()) / 1000.0
    print(f"mlp_fused level {lvl} (C={Cc}, M={M}): {len(raw)} CTAs, kernel span {rel[:, 9].max():.2f} us")
    for i, nm in enumerate(["entry", "setup", "prev_done", "a1_landed", "mma_tile0", "mma_end", "epi_first", "epi_tile0", "epi_end", "exit"]):
        col = rel[:, i][raw[:, i] > 0]
        if len(col):
            print(f"   {nm:10s} min {col.min():7.2f}  median {np.median(col):7.2f}  max {col.max():7.2f} us")
    lead = raw[raw[:, 5] > 0]
    for i, nm in enumerate(["h_full", "b1_full", "b2_full", "d1_empty", "d2_empty+a1_full", "(commit issue)"]):
        col = lead[:, 10 + i] / 1965.0
        print(f"   MMA thread waited on {nm:9s} median {np.median(col):7.2f} us  max {col.max():7.2f} us")


def main():
    h = runtime.Handle(torch.device("cuda", 0), "bf16")
    if len(sys.argv) > 2 and sys.argv[1] == "mlp":
        for lvl in sys.argv[2:]:
            mlp(h, int(lvl))
        return
    if len(sys.argv) > 2 and sys.argv[1] == "gconv":
        for lvl in sys.argv[2:]:
            gconv(h, int(lvl))
        return
    shapes = [(4096, 1536, 512), (4096, 3072, 512), (65536, 768, 128), (4096, 512, 1536), (1024, 6144, 1024), (16384, 1536, 256)]
    mode = 0
    if len(sys.argv) > 3:
        shapes = [tuple(int(v) for v in sys.argv[1:4])]
        mode = int(sys.argv[4]) if len(sys.argv) > 4 else 0     # 0 bf16 store, 2 fp32 accumulate (split-K path)
    for (M, N, K) in shapes:
        A = torch.randn(M, K, device="cuda").bfloat16(); W = torch.randn(N, K, device="cuda").bfloat16()
        bias = torch.randn(N, device="cuda"); out = torch.empty(M, N, device="cuda", dtype=torch.float32 if mode else torch.bfloat16)
        for _ in range(3):
            h.gemm(A, W, bias, out, M, N, K, out_f32=mode)
        torch.cuda.synchronize()
        h.lib.ldmb_debug_tc_trace(h.h, 1, None, 0)
        # back-to-back launches so the traced one sees a busy predecessor
        for _ in range(4):
            h.gemm(A, W, bias, out, M, N, K, out_f32=mode)
        buf = (C.c_int64 * (16 * 256))()
        n = h.lib.ldmb_debug_tc_trace(h.h, 1, buf, 256)
        h.lib.ldmb_debug_tc_trace(h.h, 0, None, 0)
        a = np.frombuffer(buf, dtype=np.int64).reshape(256, 16)[:148, :10].astype(np.float64)
        a = a[a[:, 0] > 0]
        t0 = a[:, 0].min()
        rel = (a - t0) / 1000.0
        print(f"M={M} N={N} K={K}: {len(a)} CTAs, kernel span {rel[:, 9].max():.2f} us")
        for i, nm in enumerate(NAMES):
            col = rel[:, i]
            print(f"   {nm:10s} min {col.min():7.2f}  median {np.median(col):7.2f}  max {col.max():7.2f} us")


if __name__ == "__main__":
    main()
