"""GPU parity of the DDIM sampler (ddpm.py:51-93) and of the VAE decode/encode (vae.py)."""
import pytest
import torch

from oracle import restate as R
from oracle.gen_golden import DECODER_CASES, ENCODER_CASES, UNET_CASES
from tests.gpu_util import PSNR_MIN_DB, assert_no_fault, build_decoder, build_encoder, build_unet, golden

pytestmark = pytest.mark.gpu


def _ddpm(precision):
    from ldm_image_generator_b200 import DDPM
    kw, wseed, B, H, W = UNET_CASES["unet_tiny"]
    cfg = R.UNetCfg(**kw)
    return DDPM(model=build_unet(cfg, R.make_unet_state(cfg, wseed), precision)), (B, cfg.input_channels, H, W)


def test_sampler_fp32_matches_reference_fixture():
    """Whole DDPM.sample trajectories (eval and train mode, linear and list schedules) against the reference's
    own output; x_T is the CPU-generator draw the reference made (SURVEY 8c rule 4)."""
    d, shape = _ddpm("fp32")
    fix = golden("ddim_tiny")
    for mode, training in (("eval", False), ("train", True)):
        d.train(training)
        x0 = d.sample(shape, seed=11, num_steps=8, use_autocast=False, x_T=fix["x_T"], progress=False).cpu()
        err = R.rel_l2(x0, fix["x0_" + mode])
        print("sampler fp32", mode, err)
        assert err < 2e-4, (mode, err)      # 8 steps through the 1/sqrt(abar_999)=157x amplification
    d.eval()
    x0 = d.sample(shape, seed=11, num_steps=6, use_autocast=False, schedule=[0, 40, 80, 120, 160, 199],
                  x_T=fix["x_T"], progress=False).cpu()
    assert R.rel_l2(x0, fix["x0_eval_lowt"]) < 2e-5
    assert_no_fault(d.model)


def test_sampler_bf16_conditioned_trajectory():
    d, shape = _ddpm("bf16")
    fix = golden("ddim_tiny")
    d.eval()
    x0 = d.sample(shape, seed=11, num_steps=6, schedule=[0, 40, 80, 120, 160, 199], x_T=fix["x_T"], progress=False).cpu()
    err = R.rel_l2(x0, fix["x0_eval_lowt"])
    print("sampler bf16 low-t", err)
    assert err < 2e-2


def test_sampler_torch_generator_lockstep_and_eta():
    """eta > 0: the per-step noise comes from torch's CUDA generator exactly where the reference draws it
    (one randn for x_T, one per step), so two identically seeded calls agree and the generator advances
    by the same amount whether or not eta is 0."""
    d, shape = _ddpm("fp32")
    d.eval()
    a = d.sample(shape, seed=3, num_steps=4, eta=0.5, progress=False)
    b = d.sample(shape, seed=3, num_steps=4, eta=0.5, progress=False)
    assert torch.equal(a, b)
    torch.manual_seed(9); torch.cuda.manual_seed(9)
    d.sample(shape, seed=None, num_steps=4, eta=0.0, progress=False)
    after = torch.randn(4, device="cuda")
    torch.manual_seed(9); torch.cuda.manual_seed(9)
    for _ in range(5):
        torch.randn(*shape, device="cuda")
    assert torch.equal(after, torch.randn(4, device="cuda"))
    # and eta > 0 matches the oracle's update when fed the same noises
    import random
    torch.manual_seed(3); torch.cuda.manual_seed(3)
    xT = torch.randn(*shape, device="cuda")
    noises = [torch.randn(*shape, device="cuda").cpu() for _ in range(4)]
    kw, wseed, *_ = UNET_CASES["unet_tiny"]
    cfg = R.UNetCfg(**kw)
    want = R.ddim_sample(R.make_unet_state(cfg, wseed), cfg, xT.cpu(), R.linear_steps(1000, 4), False, py_seed=3,
                         eta=0.5, noises=noises)
    assert R.rel_l2(a.cpu(), want) < 2e-4


@pytest.mark.parametrize("name", sorted(DECODER_CASES))
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_decoder_matches_reference_fixture(name, precision):
    kw, wseed, *_ = DECODER_CASES[name]
    cfg = R.DecoderCfg(**kw)
    fix = golden(name)
    dec = build_decoder(cfg, R.make_decoder_state(cfg, wseed), precision)
    with torch.no_grad():
        y = dec(fix["z"].cuda()).cpu()
        u8 = dec.decode_to_uint8(fix["z"].cuda()).cpu()
    err, db = R.rel_l2(y, fix["y"]), R.psnr(y.clamp(-1, 1), fix["y"].clamp(-1, 1))
    print(f"{name} {precision}: rel-L2 {err:.3e} PSNR {db:.1f} dB")
    assert err < (1e-5 if precision == "fp32" else 1e-2)
    assert db >= PSNR_MIN_DB
    want = R.to_uint8_image(fix["y"])
    diff = (u8.int() - want.int()).abs()
    assert int(diff.max()) <= (1 if precision == "fp32" else 3)
    if precision == "fp32":
        assert float((diff == 0).float().mean()) > 0.995      # truncation boundaries only
    assert_no_fault(dec)


@pytest.mark.parametrize("name", sorted(ENCODER_CASES))
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_encoder_matches_reference_fixture(name, precision):
    kw, wseed, *_ = ENCODER_CASES[name]
    cfg = R.EncoderCfg(**kw)
    fix = golden(name)
    enc = build_encoder(cfg, R.make_encoder_state(cfg, wseed), precision)
    with torch.no_grad():
        z = enc(fix["x"].cuda()).cpu()
    err = R.rel_l2(z, fix["y"])
    print(f"{name} {precision}: rel-L2 {err:.3e}")
    assert err < (1e-5 if precision == "fp32" else 1.5e-2)
    assert_no_fault(enc)


def test_vae_roundtrip_shapes_and_wrapper():
    from ldm_image_generator_b200 import VAE
    ec, dc = R.EncoderCfg(channels=(8, 16, 32, 64)), R.DecoderCfg(channels=(64, 32, 16, 8))
    vae = VAE(build_encoder(ec, R.make_encoder_state(ec, 5), "fp32"), build_decoder(dc, R.make_decoder_state(dc, 5), "fp32"))
    x = torch.randn(2, 3, 32, 48, device="cuda").clamp(-1, 1)
    z = vae.encode(x)
    assert tuple(z.shape) == (2, 8, 4, 6)
    y = vae.decode(z)
    assert tuple(y.shape) == (2, 3, 32, 48)
    want = R.decoder_forward(R.make_decoder_state(dc, 5), dc, R.encoder_forward(R.make_encoder_state(ec, 5), ec, x.cpu()))
    assert R.rel_l2(y.cpu(), want) < 1e-5


def test_film_precompute_chunks_and_per_step_path_agree():
    """DDPM.sample evaluates the Encodings MLP (unet.py:18-21) for all timesteps of the schedule up front, in chunks
    bounded by a memory budget.  One chunk, many small chunks and the per-step evaluation (UNet.forward outside
    sample) must give the same trajectory: bit-identical in fp32 validation mode."""
    import random
    from ldm_image_generator_b200 import _lib
    d, shape = _ddpm("fp32")
    d.eval()
    fix = golden("ddim_tiny")
    steps = [0, 30, 60, 90, 120, 150, 199]
    one = d.sample(shape, seed=5, num_steps=7, schedule=steps, x_T=fix["x_T"], progress=False)
    old = d.model.FILM_BYTES_BUDGET
    try:
        d.model.FILM_BYTES_BUDGET = 1                      # film_chunk() == 1: a precompute call per step
        assert d.model.film_chunk(shape[2], shape[3]) == 1
        many = d.sample(shape, seed=5, num_steps=7, schedule=steps, x_T=fix["x_T"], progress=False)
    finally:
        d.model.FILM_BYTES_BUDGET = old
    assert torch.equal(one, many)
    # the same updates driven step by step through the public forward (no precomputed tables)
    random.seed(5)
    x = fix["x_T"].cuda().float().clone()
    alpha = torch.cumprod(1 - d.beta, dim=0)
    for t, t_next in d.timesteps(7, steps):
        co, _ = d.ddim_scalars(alpha, int(t), int(t_next), 0)
        eps = d.model(x=x, time=torch.full((shape[0],), int(t), device="cuda"), condition=None)
        x0 = (x - co.c_eps_in * eps) / co.c_div
        x = x0 if int(t) == 0 else co.c_x0 * x0 + co.c_eps_out * eps
    assert R.rel_l2(x.cpu(), one.cpu()) < 1e-6
    assert_no_fault(d.model)


def test_full_size_decode_vs_oracle_and_shard_equivalence():
    """BASELINE configs[1] decode size: the default Decoder on a multi-image batch of 8x32x32 latents -> 256x256 RGB
    (multi-tile paths of every conv level, the halo-patch kernel at 64 channels, the fused uint8 output).  Two of the images
    against the CPU oracle; the decoder has no reductions into shared state, so a batch is bit-identical to its shards."""
    cfg = R.DecoderCfg()
    sd = R.make_decoder_state(cfg, 1234)
    dec = build_decoder(cfg, sd, "bf16")
    g = torch.Generator().manual_seed(11)
    z = torch.randn(12, 8, 32, 32, generator=g)
    with torch.no_grad():
        y = dec(z.cuda()).cpu()
        u8 = dec.decode_to_uint8(z.cuda()).cpu()
        shards = torch.cat([dec(z[i:i + 4].cuda()).cpu() for i in range(0, 12, 4)])
    assert torch.equal(shards, y)
    pick = [0, 11]
    want = R.decoder_forward(sd, cfg, z[pick])
    err, db = R.rel_l2(y[pick], want), R.psnr(y[pick].clamp(-1, 1), want.clamp(-1, 1))
    print(f"full-size decode: rel-L2 {err:.3e} PSNR {db:.1f} dB")
    assert err < 1e-2 and db >= PSNR_MIN_DB
    assert tuple(u8.shape) == (12, 256, 256, 3)
    assert int((u8[pick].int() - R.to_uint8_image(want).int()).abs().max()) <= 3
    assert_no_fault(dec)


def test_sample_independent_equals_the_batch1_loop():
    """DDPM.sample_independent(n) == the reference scripts' loop of batch-1 sample() calls (sample_ddpm.py:35-36 with
    seed=i per image; sample_ldm.py:71-72 with one seed before the loop), computed as one batch with per-image plans:
    same Python-RNG and torch-generator consumption, same images."""
    import random
    d, shape = _ddpm("fp32")
    one = (1,) + tuple(shape[1:])
    for training in (False, True):
        d.train(training)
        loop = torch.cat([d.sample(one, seed=i, num_steps=5, progress=False) for i in range(3)])
        tail_loop = (random.random(), float(torch.rand(1, device="cuda")))
        batch = d.sample_independent(3, one, seeds=[0, 1, 2], num_steps=5)
        assert tail_loop == (random.random(), float(torch.rand(1, device="cuda")))    # eta == 0: noise draws skipped by offset
        assert R.rel_l2(batch.cpu(), loop.cpu()) < 2e-5, training
        random.seed(9); torch.manual_seed(9); torch.cuda.manual_seed(9)
        loop = torch.cat([d.sample(one, seed=None, num_steps=4, eta=0.3, progress=False) for _ in range(3)])
        tail_loop = (random.random(), float(torch.rand(1, device="cuda")))
        random.seed(9); torch.manual_seed(9); torch.cuda.manual_seed(9)
        batch = d.sample_independent(3, one, seeds=None, num_steps=4, eta=0.3)
        tail_batch = (random.random(), float(torch.rand(1, device="cuda")))
        assert R.rel_l2(batch.cpu(), loop.cpu()) < 2e-5, training
        assert tail_loop == tail_batch                      # both RNG streams end in the same state
    assert_no_fault(d.model)
