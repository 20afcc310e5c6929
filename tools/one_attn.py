#!/usr/bin/env python
"""A few launches of the tcgen05 window-attention kernel at the config-2 level-0 shape (for ncu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldm_image_generator_b200 import runtime
h = runtime.Handle(torch.device("cuda", 0), "bf16")
B,H,W,C,wh,ww = 64,32,32,128,6,6
shift = int(os.environ.get("SHIFT", "0"))
qkv = torch.randn(B,H,W,3*C,device="cuda").bfloat16(); xm = torch.randn(B,H,W,C,device="cuda").bfloat16()
b_in = torch.randn(3*C,device="cuda"); out = torch.empty(B,H,W,4*C,device="cuda",dtype=torch.bfloat16)
for _ in range(4): h.window_attention(qkv,xm,b_in,out[...,3*C:],B,H,W,C,wh,ww,shift)
torch.cuda.synchronize()
assert h.device_fault() == 0
print("ok")
