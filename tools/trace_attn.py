#!/usr/bin/env python
"""Per-role cycle accounting inside window_attention_tc_kernel (ldmb_debug_tc_trace): where one lane of each warp role spends its time."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldm_image_generator_b200 import runtime
h = runtime.Handle(torch.device("cuda", 0), "bf16")
B,H,W,Cc,wh,ww,shift = 64,32,32,128,6,6,int(os.environ.get("SHIFT","3"))
qkv = torch.randn(B,H,W,3*Cc,device="cuda").bfloat16(); xm = torch.randn(B,H,W,Cc,device="cuda").bfloat16()
b_in = torch.randn(3*Cc,device="cuda"); out = torch.empty(B,H,W,4*Cc,device="cuda",dtype=torch.bfloat16)
for _ in range(3): h.window_attention(qkv,xm,b_in,out[...,3*Cc:],B,H,W,Cc,wh,ww,shift)
torch.cuda.synchronize()
h.lib.ldmb_debug_tc_trace(h.h, 1, None, 0)
for _ in range(3): h.window_attention(qkv,xm,b_in,out[...,3*Cc:],B,H,W,Cc,wh,ww,shift)
torch.cuda.synchronize()
buf = (C.c_int64 * (16 * 256))()
n = h.lib.ldmb_debug_tc_trace(h.h, 1, buf, 256)
h.lib.ldmb_debug_tc_trace(h.h, 0, None, 0)
import numpy as np
a = np.array(buf[:16*148], dtype=np.int64).reshape(148, 16)
names = ["soft:other","soft:wait_full","soft:wait_s_full","soft:ld_S+arrive","soft:wait_pv_done","soft:math","soft:epilogue","soft:P_write+fence+arrive",
         "mma:issue/other","mma:wait_full","mma:wait_s_empty","mma:wait_p_full","load:other","load:wait_empty","load:issue","load:wait_group+fence+arrive"]
items = 2304 / 148
for i, nm in enumerate(names):
    print(f"{nm:32s} median {np.median(a[:, i]):9.0f} clk  per item {np.median(a[:, i]) / items:7.0f}")
print("soft total per item", a[:, :8].sum(1).mean() / items, " mma total", a[:, 8:12].sum(1).mean() / items, " load total", a[:, 12:].sum(1).mean() / items)
