"""GPU: single kernels through the C ABI against a plain torch fp32 reference of the same op."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handles():
    from ldm_image_generator_b200 import runtime
    dev = torch.device("cuda", 0)
    return {"bf16": runtime.Handle(dev, "bf16"), "fp32": runtime.Handle(dev, "fp32")}


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


GEMM_SHAPES = [(128, 128, 64), (256, 256, 128), (1000, 384, 512), (4096, 768, 128), (130, 64, 64), (512, 96, 192),
               (16, 1024, 1024), (8192, 256, 1536), (1024, 3072, 512),
               (4096, 3072, 512), (16384, 768, 256), (4096, 1536, 512), (5000, 1280, 448),     # A-stationary tilings (bf16-out, K <= 512, > 74 pair tiles)
               (65536, 384, 128), (32768, 768, 256), (20000, 640, 192)]                       # persistent A-stationary (more m-tiles than CTA pairs, K <= 256)


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_fp32_validation_kernel(handles, M, N, K):
    h = handles["fp32"]
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g); W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    out = torch.empty(M, N, device="cuda")
    h.gemm(A, W, b, out, M, N, K)
    ref = (A.double() @ W.double().t() + b.double()).float()
    assert _rel(out, ref) < 2e-6


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("mode", ["store_bf16_relu", "store_f32", "accum_f32_leaky_none"])
def test_gemm_tcgen05_matches_fp32_reference(handles, M, N, K, mode):
    """bf16 operands, fp32 accumulate: must agree with an fp32 GEMM of the same bf16-rounded operands
    to fp32 accumulation noise (store_f32) or bf16 output rounding (store_bf16)."""
    h = handles["bf16"]
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda", generator=g)
    ref = A.double() @ W.double().t() + b.double()
    if mode == "store_bf16_relu":
        out = torch.full((M, N), 7.0, device="cuda", dtype=torch.bfloat16)
        h.gemm(A, W, b, out, M, N, K, out_f32=0, act=1)
        assert _rel(out.float(), ref.clamp_min(0)) < 4e-3
        sim = torch.empty_like(out)
        h.gemm(A, W, b, sim, M, N, K, out_f32=0, act=1, force_simt=True)
        assert _rel(out.float(), sim.float()) < 4e-3
    elif mode == "store_f32":
        out = torch.full((M, N), 7.0, device="cuda")
        h.gemm(A, W, b, out, M, N, K, out_f32=1)
        assert _rel(out, ref) < 1e-5
    else:
        base = torch.randn(M, N, device="cuda", generator=g)
        out = base.clone()
        h.gemm(A, W, b, out, M, N, K, out_f32=2)
        assert _rel(out, ref + base.double()) < 1e-5
    assert h.device_fault() == 0


CONV_SHAPES = [(2, 32, 32, 64, 64), (1, 8, 8, 128, 128), (2, 16, 16, 64, 128), (1, 128, 128, 64, 64),
               (3, 64, 64, 128, 256), (5, 4, 4, 64, 64), (1, 256, 256, 64, 64), (2, 32, 32, 512, 512)]


def _pack_conv(w):   # [N, C, 3, 3] -> [N, tap*C + c]
    N, Cc = w.shape[:2]
    return w.permute(0, 2, 3, 1).reshape(N, 9 * Cc).contiguous()


@pytest.mark.parametrize("B,H,W,C,N", CONV_SHAPES)
def test_conv3x3_tcgen05_implicit_gemm(handles, B, H, W, C, N):
    h = handles["bf16"]
    g = torch.Generator(device="cuda").manual_seed(B + H + C + N)
    x = torch.randn(B, H, W, C, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, C, 3, 3, device="cuda", generator=g) / (9 * C) ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda", generator=g)
    out = torch.full((B, H, W, N), 7.0, device="cuda", dtype=torch.bfloat16)
    h.conv3x3(x, _pack_conv(w), b, out, B, H, W, C, N, act=2)
    ref = F.leaky_relu(F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=1), 0.01).permute(0, 2, 3, 1)
    assert _rel(out.float(), ref) < 4e-3
    assert h.device_fault() == 0


@pytest.mark.parametrize("B,H,W,C,N", [(2, 12, 20, 24, 40), (1, 7, 5, 8, 8)])
def test_conv3x3_fp32_validation_kernel(handles, B, H, W, C, N):
    h = handles["fp32"]
    x = torch.randn(B, H, W, C, device="cuda"); w = torch.randn(N, C, 3, 3, device="cuda") / (9 * C) ** 0.5
    b = torch.randn(N, device="cuda")
    out = torch.empty(B, H, W, N, device="cuda")
    h.conv3x3(x, _pack_conv(w), b, out, B, H, W, C, N)
    torch.backends.cudnn.allow_tf32 = False
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double(), b.double(), padding=1).permute(0, 2, 3, 1).float()
    assert _rel(out, ref) < 2e-6


@pytest.mark.parametrize("C", [32, 128, 512, 1024, 2048])
@pytest.mark.parametrize("prec", ["bf16", "fp32"])
def test_channelnorm_film(handles, C, prec):
    """modules.py:23-25 + unet.py:22 on NHWC rows."""
    h = handles[prec]
    M, HW = 4 * 37, 37
    x = torch.randn(M, C, device="cuda") * 3 + 1
    film = torch.randn(HW, 2 * C, device="cuda")
    out = torch.empty(M, C, device="cuda", dtype=torch.bfloat16 if prec == "bf16" else torch.float32)
    h.channelnorm_film(x, film, out, M, C, HW)
    xn = (x - x.mean(1, keepdim=True)) / torch.sqrt(x.var(1, keepdim=True) + 1e-4)
    idx = torch.arange(M, device="cuda") % HW
    ref = xn * film[idx, :C] + film[idx, C:]
    assert _rel(out.float(), ref) < (4e-3 if prec == "bf16" else 2e-6)


def _window_attention_torch(qkv, C, ws):
    """Unpadded, unshifted windows: softmax(q k^T / sqrt(32)) v per (window, head), fp64."""
    B, H, W, _ = qkv.shape
    q, k, v = qkv.double().split(C, dim=-1)

    def win(t):   # [B,H,W,C] -> [B, nh, nw, heads, ws*ws, 32]
        t = t.reshape(B, H // ws, ws, W // ws, ws, C // 32, 32)
        return t.permute(0, 1, 3, 5, 2, 4, 6).reshape(B, H // ws, W // ws, C // 32, ws * ws, 32)
    s = torch.softmax(win(q) @ win(k).transpose(-1, -2) / 32 ** 0.5, dim=-1) @ win(v)
    s = s.reshape(B, H // ws, W // ws, C // 32, ws, ws, 32).permute(0, 1, 4, 2, 5, 3, 6)
    return s.reshape(B, H, W, C)


@pytest.mark.parametrize("B,H,W,C,wh,ww,shift", [(2, 12, 12, 128, 6, 6, 0), (2, 12, 12, 128, 6, 6, 3), (3, 20, 28, 256, 6, 6, 0),
                                                 (3, 20, 28, 256, 6, 6, 3), (5, 4, 4, 1024, 4, 4, 0), (2, 8, 8, 512, 6, 6, 3),
                                                 (1, 32, 32, 128, 6, 6, 3), (2, 5, 3, 128, 5, 3, 0), (64, 32, 32, 128, 6, 6, 3),
                                                 (64, 16, 16, 256, 6, 6, 0), (64, 4, 4, 1024, 4, 4, 0), (7, 6, 6, 64, 6, 6, 0)])
def test_window_attention_mma_kernel(handles, B, H, W, C, wh, ww, shift):
    """tcgen05 (and mma.sync) attention core vs the CUDA-core kernel (validated against the reference through the UNet
    fixtures), including zero-padded windows, the rolled float key bias and the global small-image case."""
    h = handles["bf16"]
    g = torch.Generator(device="cuda").manual_seed(B * 131 + H * 17 + C + shift)
    qkv = torch.randn(B, H, W, 3 * C, device="cuda", generator=g).bfloat16()
    xm = torch.randn(B, H, W, C, device="cuda", generator=g).bfloat16()
    b_in = torch.randn(3 * C, device="cuda", generator=g) * 0.5
    ldo = 4 * C
    out = torch.full((B, H, W, ldo), 7.0, device="cuda", dtype=torch.bfloat16)
    ref = torch.full((B, H, W, ldo), 7.0, device="cuda", dtype=torch.bfloat16)
    h.window_attention(qkv, xm, b_in, out[..., 3 * C:], B, H, W, C, wh, ww, shift)                       # tcgen05 kernel
    h.window_attention(qkv, xm, b_in, ref[..., 3 * C:], B, H, W, C, wh, ww, shift, force_simt=1)         # CUDA-core kernel
    assert torch.equal(out[..., :3 * C], ref[..., :3 * C])          # nothing outside the output columns is touched
    assert _rel(out[..., 3 * C:].float(), ref[..., 3 * C:].float()) < 6e-3
    mma = torch.full((B, H, W, ldo), 7.0, device="cuda", dtype=torch.bfloat16)
    h.window_attention(qkv, xm, b_in, mma[..., 3 * C:], B, H, W, C, wh, ww, shift, force_simt=2)         # mma.sync kernel
    assert _rel(mma[..., 3 * C:].float(), ref[..., 3 * C:].float()) < 6e-3
    if shift == 0 and H % wh == 0 and W % ww == 0:
        want = _window_attention_torch(qkv, C, wh) if wh == ww else None
        if want is not None:
            assert _rel(out[..., 3 * C:].float(), want) < 6e-3
    assert h.device_fault() == 0


@pytest.mark.parametrize("B,H,W,C,shift", [(2, 12, 12, 128, 0), (2, 12, 12, 128, 3), (2, 20, 28, 256, 3), (2, 20, 28, 256, 0),
                                           (2, 8, 8, 512, 3), (2, 8, 8, 512, 0), (1, 32, 32, 128, 3), (1, 32, 32, 128, 0),
                                           (3, 4, 4, 1024, 0), (2, 16, 16, 256, 3), (1, 10, 7, 128, 3)])
def test_window_attention_core_vs_oracle(handles, B, H, W, C, shift):
    """The attention core against oracle.restate.window_attention DIRECTLY (attention.py:13-85 + torch MHA), shifted and
    zero-padded windows included: in-projection done in torch (the kernel's input is qkv), out-projection = identity."""
    from oracle import restate as R
    h = handles["bf16"]
    g = torch.Generator(device="cuda").manual_seed(B * 31 + H * 7 + W + C + shift)
    xm = torch.randn(B, H, W, C, device="cuda", generator=g).bfloat16()
    w_in = (torch.randn(3 * C, C, device="cuda", generator=g) / C ** 0.5).bfloat16().float()
    b_in = torch.randn(3 * C, device="cuda", generator=g) * 0.5
    qkv = (xm.float().reshape(-1, C) @ w_in.t() + b_in).reshape(B, H, W, 3 * C).bfloat16()
    out = torch.full((B, H, W, C), 7.0, device="cuda", dtype=torch.bfloat16)
    glob = H <= 6 and W <= 6
    h.window_attention(qkv, xm, b_in, out, B, H, W, C, H if glob else 6, W if glob else 6, 0 if glob else shift)
    sd = {"a.attention.in_proj_weight": w_in.cpu(), "a.attention.in_proj_bias": b_in.cpu(),
          "a.attention.out_proj.weight": torch.eye(C), "a.attention.out_proj.bias": torch.zeros(C)}
    want = R.window_attention(sd, "a.", xm.float().cpu().permute(0, 3, 1, 2).contiguous(), shift).permute(0, 2, 3, 1)
    err = _rel(out.float().cpu(), want)
    print("window attention vs oracle", (B, H, W, C, shift), err)
    assert err < 8e-3          # qkv and the output are rounded to bf16 on the kernel side only
    assert h.device_fault() == 0


def _pack_grouped(w):
    """[C, 32, 3, 3] grouped-conv weight -> block-diagonal pairs [C/64][64][9*64] (ldmb.h: ldmb_grouped_conv3x3)."""
    C = w.shape[0]
    out = torch.zeros(C // 64, 64, 9, 64, device=w.device, dtype=w.dtype)
    wp = w.reshape(C // 64, 2, 32, 32, 9)                       # pair, group-in-pair, co, ci, tap
    for gl in range(2):
        out[:, gl * 32:(gl + 1) * 32, :, gl * 32:(gl + 1) * 32] = wp[:, gl].permute(0, 1, 3, 2)
    return out.reshape(C, 9 * 64).contiguous()


@pytest.mark.parametrize("B,H,W,C", [(2, 32, 32, 128), (3, 16, 16, 256), (5, 8, 8, 512), (7, 4, 4, 1024), (2, 20, 28, 128),
                                     (1, 64, 64, 256), (3, 5, 3, 64), (64, 8, 8, 512)])
@pytest.mark.parametrize("generic", [False, True])
def test_grouped_conv3x3_accumulates_into_residual(handles, B, H, W, C, generic):
    """unet.py:30 on NHWC: halo-patch tcgen05 kernel (and the generic 9-tap-load kernel) vs F.conv2d(groups=C/32)."""
    h = handles["bf16"]
    g = torch.Generator(device="cuda").manual_seed(B * 7 + H * 3 + W + C)
    xm = torch.randn(B, H, W, C, device="cuda", generator=g).bfloat16()
    w = (torch.randn(C, 32, 3, 3, device="cuda", generator=g) / 288 ** 0.5).bfloat16()
    b = torch.randn(C, device="cuda", generator=g)
    x0 = torch.randn(B, H, W, C, device="cuda", generator=g)
    x = x0.clone()
    h.grouped_conv3x3(xm, _pack_grouped(w), b, x, B, H, W, C, force_generic=generic)
    ref = F.conv2d(xm.float().permute(0, 3, 1, 2), w.float(), b, padding=1, groups=C // 32).permute(0, 2, 3, 1)
    assert h.device_fault() == 0
    assert _rel(x - x0, ref) < 1e-5


@pytest.mark.parametrize("B,H,W,C", [(64, 8, 8, 512), (64, 4, 4, 1024), (3, 8, 8, 512), (7, 4, 4, 1024), (5, 8, 8, 256), (130, 4, 4, 512),
                                     (2, 8, 8, 128), (9, 4, 4, 64), (300, 8, 8, 512), (1, 2, 2, 1024)])
def test_normconv_fused_kernel(handles, B, H, W, C):
    """ChannelNorm + FiLM + grouped conv in one kernel (cluster statistics exchange, in-place residual update) vs torch:
    xm against the fp32 formula (bf16 rounding), the residual update against F.conv2d of the kernel's own bf16 xm."""
    h = handles["bf16"]
    g = torch.Generator(device="cuda").manual_seed(B * 11 + H * 5 + C)
    x0 = torch.randn(B, H, W, C, device="cuda", generator=g) * 2 + 0.5
    film = torch.randn(H * W, 2 * C, device="cuda", generator=g)
    w = (torch.randn(C, 32, 3, 3, device="cuda", generator=g) / 288 ** 0.5).bfloat16()
    b = torch.randn(C, device="cuda", generator=g)
    x = x0.clone()
    xm = torch.full((B, H, W, C), 7.0, device="cuda", dtype=torch.bfloat16)
    h.normconv(x, film, xm, _pack_grouped(w), b, B, H, W, C)
    assert h.device_fault() == 0
    xn = (x0 - x0.mean(-1, keepdim=True)) / torch.sqrt(x0.var(-1, keepdim=True) + 1e-4)
    f = film.reshape(1, H, W, 2 * C)
    want_xm = xn * f[..., :C] + f[..., C:]
    assert _rel(xm.float(), want_xm) < 4e-3
    ref = F.conv2d(xm.float().permute(0, 3, 1, 2), w.float(), b, padding=1, groups=C // 32).permute(0, 2, 3, 1)
    assert _rel(x - x0, ref) < 1e-5
    # and the separate kernels produce the same xm bit for bit up to the statistics' summation order
    xm2 = torch.empty_like(xm)
    h.channelnorm_film(x0.reshape(-1, C), film, xm2.reshape(-1, C), B * H * W, C, H * W)
    assert _rel(xm.float(), xm2.float()) < 2e-3


@pytest.mark.parametrize("M,C,e1,e2", [(256, 128, 0, 1), (4096, 128, 3, 2), (1000, 128, 1, 3), (65536, 128, 2, 0), (512, 256, 0, 3),
                                       (16384, 256, 3, 1), (700, 256, 2, 1),
                                       (256, 512, 0, 1), (4096, 512, 1, 2), (700, 512, 3, 0), (8192, 512, 2, 3)])     # C = 512: the 8-CTA-cluster kernel
def test_mlp_fused_kernel(handles, M, C, e1, e2):
    """Fused ReGLU feed-forward (general + 2 picked experts) vs fp64 torch on the same bf16-rounded operands."""
    h = handles["bf16"]
    g = torch.Generator(device="cuda").manual_seed(M + C + e1 * 5 + e2)
    xm = torch.randn(M, C, device="cuda", generator=g).bfloat16()
    wa = (torch.randn(5, C, C, device="cuda", generator=g) / C ** 0.5).bfloat16()
    wb = (torch.randn(5, C, C, device="cuda", generator=g) / C ** 0.5).bfloat16()
    wc = (torch.randn(5, C, C, device="cuda", generator=g) / C ** 0.5).bfloat16()
    ba, bb, bc = (torch.randn(5, C, device="cuda", generator=g) * 0.3 for _ in range(3))
    # a|b rows interleaved in chunks of 64 per expert (ldmb.h)
    w_ab = torch.stack([wa.reshape(5, C // 64, 64, C), wb.reshape(5, C // 64, 64, C)], dim=2).reshape(5 * 2 * C, C).contiguous()
    b_ab = torch.stack([ba.reshape(5, C // 64, 64), bb.reshape(5, C // 64, 64)], dim=2).reshape(5 * 2 * C).contiguous()
    x0 = torch.randn(M, C, device="cuda", generator=g)
    x = x0.clone()
    h.mlp_fused(xm, w_ab, b_ab, wc.reshape(5 * C, C).contiguous(), bc.reshape(5 * C).contiguous(), x, M, C, e1, e2)
    ref = torch.zeros(M, C, device="cuda", dtype=torch.float64)
    for e in (0, 1 + e1, 1 + e2):
        hh = (xm.double() @ wa[e].double().t() + ba[e].double()) * torch.relu(xm.double() @ wb[e].double().t() + bb[e].double())
        ref += hh.bfloat16().double() @ wc[e].double().t() + bc[e].double()       # h is rounded to bf16 between the two GEMMs
    assert h.device_fault() == 0
    assert _rel(x - x0, ref) < 2e-3


@pytest.mark.parametrize("M,C,e1,e2", [(256, 128, 0, 1), (65536, 128, 2, 3), (1000, 128, 3, 1), (512, 256, 1, 0), (16384, 256, 2, 3), (700, 256, 0, 2)])
def test_mlp_fused_with_out_proj(handles, M, C, e1, e2):
    """Attention blocks: the fused feed-forward kernel also adds the MHA out_proj of the attention output (extra K-chunks of the
    c-projection accumulator) -- vs fp64 torch on the same bf16-rounded operands; att is a strided view like the step's hbuf."""
    h = handles["bf16"]
    g = torch.Generator(device="cuda").manual_seed(M + C + e1 * 5 + e2 + 17)
    xm = torch.randn(M, C, device="cuda", generator=g).bfloat16()
    wa = (torch.randn(5, C, C, device="cuda", generator=g) / C ** 0.5).bfloat16()
    wb = (torch.randn(5, C, C, device="cuda", generator=g) / C ** 0.5).bfloat16()
    wc = (torch.randn(6, C, C, device="cuda", generator=g) / C ** 0.5).bfloat16()        # slot 5 = out_proj.weight
    ba, bb = (torch.randn(5, C, device="cuda", generator=g) * 0.3 for _ in range(2))
    bc = torch.randn(6, C, device="cuda", generator=g) * 0.3
    hbuf = torch.randn(M, 4 * C, device="cuda", generator=g).bfloat16()
    att = hbuf[:, 3 * C:]
    w_ab = torch.stack([wa.reshape(5, C // 64, 64, C), wb.reshape(5, C // 64, 64, C)], dim=2).reshape(5 * 2 * C, C).contiguous()
    b_ab = torch.stack([ba.reshape(5, C // 64, 64), bb.reshape(5, C // 64, 64)], dim=2).reshape(5 * 2 * C).contiguous()
    x0 = torch.randn(M, C, device="cuda", generator=g)
    x = x0.clone()
    w_c, b_c = wc.reshape(6 * C, C).contiguous(), bc.reshape(6 * C).contiguous()
    h.mlp_fused_attn(xm, w_ab, b_ab, w_c, b_c, att, x, M, C, e1, e2)
    ref = att.double() @ wc[5].double().t() + bc[5].double()
    for e in (0, 1 + e1, 1 + e2):
        hh = (xm.double() @ wa[e].double().t() + ba[e].double()) * torch.relu(xm.double() @ wb[e].double().t() + bb[e].double())
        ref += hh.bfloat16().double() @ wc[e].double().t() + bc[e].double()
    torch.cuda.synchronize()
    assert h.device_fault() == 0
    assert _rel(x - x0, ref) < 2e-3
