#!/usr/bin/env python
"""%globaltimer stamps at the phase boundaries of normconv_kernel (ldmb_debug_tc_trace)."""
import ctypes as C, os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldm_image_generator_b200 import runtime
h = runtime.Handle(torch.device("cuda", 0), "bf16")
B, H, W, Cc = [int(v) for v in os.environ.get("SHAPE", "64,8,8,512").split(",")]
x = torch.randn(B, H, W, Cc, device="cuda"); film = torch.randn(H * W, 2 * Cc, device="cuda")
w = torch.zeros(Cc, 9 * 64, device="cuda", dtype=torch.bfloat16); b = torch.randn(Cc, device="cuda")
xm = torch.empty(B, H, W, Cc, device="cuda", dtype=torch.bfloat16)
for _ in range(3): h.normconv(x, film, xm, w, b, B, H, W, Cc)
torch.cuda.synchronize()
h.lib.ldmb_debug_tc_trace(h.h, 1, None, 0)
for _ in range(2): h.normconv(x, film, xm, w, b, B, H, W, Cc)
torch.cuda.synchronize()
buf = (C.c_int64 * (16 * 256))()
h.lib.ldmb_debug_tc_trace(h.h, 1, buf, 256)
h.lib.ldmb_debug_tc_trace(h.h, 0, None, 0)
a = np.array(buf[:16 * 256], dtype=np.int64).reshape(256, 16)
a = a[a[:, 0] > 0]
t0 = a[:, 0].min()
names = ["entry", "init done", "cluster up", "prev kernel done", "x loaded+partials sent", "cluster barrier", "phase B done", "patch sync", "mma done", "round end", "exit"]
print(f"{len(a)} CTAs")
for i, n in enumerate(names):
    c = (a[:, i] - t0) / 1e3
    print(f"  {n:26s} min {c.min():7.2f}  median {np.median(c):7.2f}  max {c.max():7.2f} us")
