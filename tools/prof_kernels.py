#!/usr/bin/env python
"""Launch single kernels of the config-2 workload eagerly through the C ABI (for `ncu --set full -k regex:...`).
WHICH=gconv:2 | gemm_ab:3 | gemm_c:2 | norm:0 | attn:0 | mlp:0 (kernel:level); each is launched REPS times."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ldm_image_generator_b200 import runtime  # noqa: E402

which, lvl = os.environ.get("WHICH", "gconv:2").split(":")
lvl = int(lvl)
reps = int(os.environ.get("REPS", "4"))
B = 64
C, H = 128 << lvl, 32 >> lvl
M = B * H * H
h = runtime.Handle(torch.device("cuda", 0), "bf16")
g = torch.Generator(device="cuda").manual_seed(0)
xm = torch.randn(B, H, H, C, device="cuda", generator=g).bfloat16()
x = torch.randn(B, H, H, C, device="cuda", generator=g)
for _ in range(reps):
    if which == "gconv":
        w = torch.randn(C, 576, device="cuda", generator=g).bfloat16()
        h.grouped_conv3x3(xm, w, torch.zeros(C, device="cuda"), x, B, H, H, C)
    elif which == "gemm_ab":
        w = torch.randn(6 * C, C, device="cuda", generator=g).bfloat16()
        out = torch.empty(M, 6 * C, device="cuda", dtype=torch.bfloat16)
        h.gemm(xm, w, torch.zeros(6 * C, device="cuda"), out, M, 6 * C, C)
    elif which == "gemm_c":
        a = torch.randn(M, 3 * C, device="cuda", generator=g).bfloat16()
        w = torch.randn(C, 3 * C, device="cuda", generator=g).bfloat16()
        h.gemm(a, w, torch.zeros(C, device="cuda"), x, M, C, 3 * C, out_f32=2)
    elif which == "norm":
        film = torch.randn(H * H, 2 * C, device="cuda", generator=g)
        h.channelnorm_film(x, film, xm, M, C, H * H)
    elif which == "mlp":
        w_ab = torch.randn(10 * C, C, device="cuda", generator=g).bfloat16(); w_c = torch.randn(5 * C, C, device="cuda", generator=g).bfloat16()
        h.mlp_fused(xm, w_ab, torch.zeros(10 * C, device="cuda"), w_c, torch.zeros(5 * C, device="cuda"), x, M, C, 1, 2)
    elif which == "attn":
        qkv = torch.randn(B, H, H, 3 * C, device="cuda", generator=g).bfloat16()
        att = torch.empty(B, H, H, 4 * C, device="cuda", dtype=torch.bfloat16)
        ws = min(6, H)
        h.window_attention(qkv, xm, torch.zeros(3 * C, device="cuda"), att[..., 3 * C:], B, H, H, C, ws, ws, 3 if H > 6 else 0)
torch.cuda.synchronize()
assert h.device_fault() == 0
print("ok", which, lvl)
