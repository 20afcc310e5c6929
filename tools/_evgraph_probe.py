import torch
x = torch.randn(4096, 4096, device="cuda")
s = torch.cuda.Stream()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
with torch.cuda.stream(s):
    y = x @ x
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g, stream=s):
        evs[0].record()
        y = x @ x
        evs[1].record()
        z = y @ x
        evs[2].record()
    g.replay(); torch.cuda.synchronize()
    print("elapsed in graph:", evs[0].elapsed_time(evs[1]), evs[1].elapsed_time(evs[2]))
except Exception as e:
    print("FAILED", type(e).__name__, e)
