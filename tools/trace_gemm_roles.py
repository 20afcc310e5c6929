#!/usr/bin/env python
"""Per-role cycle accounting inside gemm_tc_kernel (a library built with LDMB_EXTRA_NVCC_FLAGS=-DLDMB_TC_TRACE, loaded through
LDMB_LIB_PATH): where the epilogue warps and the MMA warp of the tcgen05 GEMM spend their cycles.   trace_gemm_roles.py M N K [mode]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ldm_image_generator_b200 import runtime
h = runtime.Handle(torch.device("cuda", 0), "bf16")
shapes = [(65536, 384, 128), (16384, 768, 256), (4096, 3072, 512)]
mode = 0
if len(sys.argv) > 3:
    shapes = [tuple(int(v) for v in sys.argv[1:4])]
    mode = int(sys.argv[4]) if len(sys.argv) > 4 else 0          # 0 bf16 store, 2 fp32 accumulate (residual GEMMs: split-K, TMA reduce-add)
for (M, N, K) in shapes:
    A = torch.randn(M, K, device="cuda").bfloat16(); W = torch.randn(N, K, device="cuda").bfloat16()
    bias = torch.randn(N, device="cuda"); out = torch.zeros(M, N, device="cuda", dtype=torch.float32 if mode else torch.bfloat16)
    for _ in range(3): h.gemm(A, W, bias, out, M, N, K, out_f32=mode)
    torch.cuda.synchronize()
    h.lib.ldmb_debug_tc_trace(h.h, 1, None, 0)
    for _ in range(4): h.gemm(A, W, bias, out, M, N, K, out_f32=mode)
    buf = (C.c_int64 * (16 * 512))()
    h.lib.ldmb_debug_tc_trace(h.h, 1, buf, 512)
    h.lib.ldmb_debug_tc_trace(h.h, 0, None, 0)
    a = np.array(buf[:], dtype=np.int64).reshape(512, 16)
    st, acc = a[:256].astype(np.float64), a[256:].astype(np.float64)
    live = st[:, 0] > 0
    span = (st[live, 9].max() - st[live, 0].min()) / 1e3
    print(f"M={M} N={N} K={K}: {int(live.sum())} CTAs, kernel span {span:.2f} us")
    e = acc[live]
    for i, nm in enumerate(["epi: other / loop", "epi: bias staging + bar.sync", "epi: wait accumulator (tfull)", "epi: wait slab free (bulk wait) + syncwarp",
                            "epi: (unused)", "epi: tcgen05.ld + math + st.shared", "epi: fence + TMA store issue", "epi: fences + arrive tempty",
                            "mma: wait accumulator free (tempty)", "mma: wait operands (full)", "mma: issue + other"]):
        col = e[:, i][e[:, i] > 0] if (e[:, i] > 0).any() else np.zeros(1)
        print(f"   {nm:44s} median {np.median(col) / 1965.0:7.2f} us")
