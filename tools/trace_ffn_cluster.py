#!/usr/bin/env python
"""%globaltimer stamps of ffn_cluster_kernel (C = 512) per CTA: when each cluster starts / its first operands land / ends."""
import ctypes as C, os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldm_image_generator_b200 import runtime
h = runtime.Handle(torch.device("cuda", 0), "bf16")
M, Cc = int(os.environ.get("M", "4096")), 512
xm = torch.randn(M, Cc, device="cuda").bfloat16()
w_ab = (torch.randn(10 * Cc, Cc, device="cuda") / Cc ** 0.5).bfloat16(); b_ab = torch.randn(10 * Cc, device="cuda")
w_c = (torch.randn(5 * Cc, Cc, device="cuda") / Cc ** 0.5).bfloat16(); b_c = torch.randn(5 * Cc, device="cuda")
x = torch.zeros(M, Cc, device="cuda")
for _ in range(3): h.mlp_fused(xm, w_ab, b_ab, w_c, b_c, x, M, Cc, 1, 2)
torch.cuda.synchronize()
h.lib.ldmb_debug_tc_trace(h.h, 1, None, 0)
h.mlp_fused(xm, w_ab, b_ab, w_c, b_c, x, M, Cc, 1, 2)
torch.cuda.synchronize()
buf = (C.c_int64 * (16 * 512))()
h.lib.ldmb_debug_tc_trace(h.h, 1, buf, 512)
h.lib.ldmb_debug_tc_trace(h.h, 0, None, 0)
grid = min(256, (M + 255) // 256 * 8)
a = np.array(buf[:], dtype=np.int64).reshape(512, 16)[:grid].astype(np.float64)
t0 = a[:, 0].min()
names = {0: "entry", 1: "setup", 2: "xm issue", 3: "first operands", 5: "mma issued", 6: "first D1", 7: "D2 full", 8: "epi end", 9: "exit"}
print(f"M={M}: {grid} CTAs, span {(a[:, 9].max() - t0) / 1e3:.2f} us")
for cl in range(grid // 8):
    c = a[cl * 8:(cl + 1) * 8]
    lead = c[0]
    print(f"cluster {cl:2d}: " + "  ".join(f"{names[k]} {(lead[k] - t0) / 1e3:6.2f}" for k in (0, 2, 3, 6, 5, 7, 9)) + f"   | exit max {(c[:, 9].max() - t0) / 1e3:6.2f}")
