import sys, os, torch
sys.path.insert(0, os.getcwd())
from ldm_image_generator_b200 import runtime
h = runtime.Handle(torch.device("cuda", 0), "bf16")
for (B,H,W,C,wh,ww,shift) in [(64,32,32,128,6,6,3),(64,32,32,128,6,6,0),(64,16,16,256,6,6,3),(64,8,8,512,6,6,3),(64,4,4,1024,4,4,0)]:
    qkv = torch.randn(B,H,W,3*C,device="cuda").bfloat16(); xm = torch.randn(B,H,W,C,device="cuda").bfloat16()
    b_in = torch.randn(3*C,device="cuda"); out = torch.empty(B,H,W,4*C,device="cuda",dtype=torch.bfloat16)
    for mode in (0,2):
        for _ in range(3): h.window_attention(qkv,xm,b_in,out[...,3*C:],B,H,W,C,wh,ww,shift,force_simt=mode)
        torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): h.window_attention(qkv,xm,b_in,out[...,3*C:],B,H,W,C,wh,ww,shift,force_simt=mode)
        e1.record(); torch.cuda.synchronize()
        print(f"attn B={B} {H}x{W} C={C} shift={shift} mode={'tcgen05' if mode==0 else 'mma.sync'}: {e0.elapsed_time(e1)/20*1e3:8.1f} us", flush=True)
    assert h.device_fault()==0
