#!/usr/bin/env python
"""Window-attention core timings (tcgen05 vs mma.sync) at the config-2 shapes; LDMB_ATTN_DBG ablation bits apply to mode 0."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldm_image_generator_b200 import runtime
h = runtime.Handle(torch.device("cuda", 0), "bf16")
modes = [int(m) for m in os.environ.get("MODES", "0,2").split(",")]
shapes = [(64,32,32,128,6,6,3),(64,32,32,128,6,6,0),(64,16,16,256,6,6,3),(64,8,8,512,6,6,3),(64,4,4,1024,4,4,0)]
if os.environ.get("SHAPES"): shapes = shapes[:int(os.environ["SHAPES"])]
for (B,H,W,C,wh,ww,shift) in shapes:
    qkv = torch.randn(B,H,W,3*C,device="cuda").bfloat16(); xm = torch.randn(B,H,W,C,device="cuda").bfloat16()
    b_in = torch.randn(3*C,device="cuda"); out = torch.empty(B,H,W,4*C,device="cuda",dtype=torch.bfloat16)
    for mode in modes:
        f = lambda: h.window_attention(qkv,xm,b_in,out[...,3*C:],B,H,W,C,wh,ww,shift,force_simt=mode)
        for _ in range(3): f()
        g = torch.cuda.CUDAGraph()          # graph replay: device time without the per-call host overhead
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            f()
            with torch.cuda.graph(g, stream=s):
                for _ in range(20): f()
        torch.cuda.synchronize()
        g.replay(); torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        print(f"attn B={B} {H}x{W} C={C} shift={shift} {'tcgen05 ' if mode==0 else 'mma.sync'}: {e0.elapsed_time(e1)/20*1e3:8.1f} us", flush=True)
    assert h.device_fault()==0
