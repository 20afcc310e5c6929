#!/usr/bin/env python
"""Step-by-step timeline inside ffn_cluster_kernel (C = 512): %globaltimer stamps of the leader MMA warp (start of every schedule
step), the non-leader's relay warp and the gate epilogue (warp 2) of one cluster, through the LDMB_FFN_PROGRESS debug buffer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
M, Cc = int(os.environ.get("M", "4096")), 512
grid = (M + 255) // 256 * 8
prog = torch.zeros(grid * 64, dtype=torch.int32, device="cuda")
os.environ["LDMB_FFN_PROGRESS"] = str(prog.data_ptr())
from ldm_image_generator_b200 import runtime
h = runtime.Handle(torch.device("cuda", 0), "bf16")
xm = torch.randn(M, Cc, device="cuda").bfloat16()
w_ab = (torch.randn(10 * Cc, Cc, device="cuda") / Cc ** 0.5).bfloat16(); b_ab = torch.randn(10 * Cc, device="cuda")
w_c = (torch.randn(5 * Cc, Cc, device="cuda") / Cc ** 0.5).bfloat16(); b_c = torch.randn(5 * Cc, device="cuda")
x = torch.zeros(M, Cc, device="cuda")
for _ in range(3): h.mlp_fused(xm, w_ab, b_ab, w_c, b_c, x, M, Cc, 1, 2)
torch.cuda.synchronize()
p = prog.cpu().numpy().astype("int64").reshape(grid, 64)
cl = int(os.environ.get("CLUSTER", "0"))
c = p[cl * 8:(cl + 1) * 8]
t0 = c[c != 0].min()
us = lambda v: (v - t0) / 1e3
sched = os.environ.get("LDMB_FFN_SCHED", "0")
for pr in range(4):
    lead, rel = c[2 * pr], c[2 * pr + 1]
    print(f"pair {pr} leader MMA step starts (us): " + " ".join(f"{us(lead[s]):5.2f}" for s in range(31)))
    print(f"pair {pr} relay  G2 step starts (us):  " + " ".join(f"{us(rel[s]):5.2f}" if rel[s] else "  -  " for s in range(30)))
    for who, row in (("leader", lead), ("non-leader", rel)):
        print(f"pair {pr} {who} gate (d1_full, slot free, written, copies issued) per chunk: " +
              " | ".join(" ".join(f"{us(row[32 + 4 * k + j]):5.2f}" for j in range(4)) for k in range(6)))
