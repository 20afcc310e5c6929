#!/usr/bin/env python
"""Per-role cycle accounting inside mlp_fused_kernel (build with LDMB_EXTRA_NVCC_FLAGS=-DLDMB_MLP_TRACE)."""
import ctypes as C, os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldm_image_generator_b200 import runtime
h = runtime.Handle(torch.device("cuda", 0), "bf16")
for (M, Cc) in [(65536, 128), (16384, 256)]:
    xm = torch.randn(M, Cc, device="cuda").bfloat16()
    w_ab = (torch.randn(10 * Cc, Cc, device="cuda") / Cc ** 0.5).bfloat16(); b_ab = torch.randn(10 * Cc, device="cuda")
    w_c = (torch.randn(5 * Cc, Cc, device="cuda") / Cc ** 0.5).bfloat16(); b_c = torch.randn(5 * Cc, device="cuda")
    x = torch.zeros(M, Cc, device="cuda")
    for _ in range(3): h.mlp_fused(xm, w_ab, b_ab, w_c, b_c, x, M, Cc, 1, 2)
    torch.cuda.synchronize()
    h.lib.ldmb_debug_tc_trace(h.h, 1, None, 0)
    for _ in range(3): h.mlp_fused(xm, w_ab, b_ab, w_c, b_c, x, M, Cc, 1, 2)
    torch.cuda.synchronize()
    buf = (C.c_int64 * (16 * 512))()
    h.lib.ldmb_debug_tc_trace(h.h, 1, buf, 512)
    h.lib.ldmb_debug_tc_trace(h.h, 0, None, 0)
    a = np.array(buf[:], dtype=np.int64).reshape(512, 16)[256:]
    grid = 148 if M // 256 >= 74 else (M // 256) * 2
    lead = a[0:grid:2]          # leader CTAs hold the MMA accounting
    units = (M / 256) / (grid / 2) * 6 * (Cc // 128)
    print(f"mlp_fused C={Cc} M={M}: grid {grid}, {units:.1f} units per CTA pair")
    for i, nm in enumerate(["mma:issue/other", "mma:wait d1_empty", "mma:wait w1_full", "mma:wait h_full", "mma:wait w2_full", "mma:wait a1_full", "mma:wait d2_empty"]):
        print(f"   {nm:28s} median {np.median(lead[:, i]):9.0f} clk   per unit {np.median(lead[:, i]) / units:7.0f}")
    e = a[:grid]
    for i, nm in enumerate(["epi:other", "epi:wait g1_done", "epi:wait g2_done(h slot)", "epi:ld+gate+st", "epi:fence+arrive", "epi:wait d2_full", "epi:epilogue2+barrier"]):
        print(f"   {nm:28s} median {np.median(e[:, 8 + i]):9.0f} clk   per unit {np.median(e[:, 8 + i]) / units:7.0f}")
