"""Parity at the sizes of BASELINE.json's other configurations (bench.py measures configs[1] only; these are the
parity-test cases): configs[2] = 512 images sharded by image, configs[3] = the wide UNet (2x base channels, 1.53 B
parameters) on 64x64 latents, configs[4] = the VAE encode+decode round trip at 512x512.  The CPU oracle checks a few images
at full model size (it finishes in seconds at batch 1); the rest of each batch is covered by the property the sharded
runs rely on: a batch equals the concatenation of its shards."""
import random

import pytest
import torch

from oracle import restate as R
from tests.gpu_util import BF16_STEP_TOL, PSNR_MIN_DB, assert_no_fault, build_decoder, build_encoder, build_unet

pytestmark = pytest.mark.gpu


def _run(model, x, t, plan):
    with torch.no_grad():
        return model._run(x.cuda(), [int(v) for v in t], plan=plan).cpu()


def test_config2_batch512_equals_its_shards_and_the_oracle():
    """configs[2]: 512 images at latent 32x32 through the default UNet, one GPU holding the whole batch (the 1-GPU point of
    the scaling sweep) against the 256 / 128 / 64-image shards the 2 / 4 / 8-GPU runs hold, and against the oracle."""
    from ldm_image_generator_b200 import DDPM
    cfg = R.UNetCfg()
    sd = R.make_unet_state(cfg, 1234)
    model = build_unet(cfg, sd, "bf16")
    g = torch.Generator().manual_seed(21)
    x = torch.randn(512, 8, 32, 32, generator=g)
    random.seed(9)
    plan = R.draw_plan(len(R.block_table(cfg)), False)
    t = [611] * 512
    model.set_deterministic(True)              # bit-reproducible mode: shards must be EXACTLY the batch
    y = _run(model, x, t, plan)
    for per_gpu in (256, 64):
        shards = torch.cat([_run(model, x[i:i + per_gpu], t[:per_gpu], plan) for i in range(0, 512, per_gpu)])
        assert torch.equal(shards, y), per_gpu
    pick = [0, 255, 256, 511]
    want = R.unet_forward(sd, cfg, x[pick], torch.tensor([611] * len(pick)), plan)
    err = R.rel_l2(y[pick], want)
    print(f"configs[2] B=512 step vs oracle: rel-L2 {err:.3e}")
    assert err < BF16_STEP_TOL, err
    model.set_deterministic(False)
    assert R.rel_l2(_run(model, x, t, plan), y) < 3e-3
    # the sampler over the whole batch (FiLM tables for every step precomputed once) against per-shard sampler runs
    model.set_deterministic(True)
    d = DDPM(model=model)
    d.eval()
    kw = dict(seed=4, num_steps=3, progress=False)
    whole = d.sample((512, 8, 32, 32), x_T=x, **kw).cpu()
    parts = torch.cat([d.sample((128, 8, 32, 32), x_T=x[i:i + 128], **kw).cpu() for i in range(0, 512, 128)])
    assert torch.equal(whole, parts)
    assert torch.isfinite(whole).all()
    assert_no_fault(model)


def test_config3_wide_unet_latent64_vs_oracle():
    """configs[3]: UNet(channels=[256,512,1024,2048]) on 8x64x64 latents (512x512 images): C = 2048 GEMMs with K up to 8192,
    32 heads x 2 at the deepest level, 11x11 windows of 6x6 tokens with padding at level 0.  One train-mode step of a
    3-image batch with per-sample timesteps against the oracle, then the 16-images-per-GPU batch of the config against its
    shards, then three sampler steps (one of them the t = 0 branch) against the oracle's DDIM loop."""
    from ldm_image_generator_b200 import DDPM
    cfg = R.UNetCfg(channels=(256, 512, 1024, 2048))
    sd = R.make_unet_state(cfg, 4321)
    model = build_unet(cfg, sd, "bf16")
    g = torch.Generator().manual_seed(3)
    x = torch.randn(16, 8, 64, 64, generator=g)
    random.seed(12)
    plan = R.draw_plan(len(R.block_table(cfg)), True)
    assert any(p[0] for p in plan)
    t3 = [999, 17, 480]
    y3 = _run(model, x[:3], t3, plan)
    want = R.unet_forward(sd, cfg, x[:3], torch.tensor(t3), plan)
    err = R.rel_l2(y3, want)
    print(f"configs[3] wide UNet step vs oracle: rel-L2 {err:.3e}")
    assert err < BF16_STEP_TOL, err
    model.set_deterministic(True)
    t = [480] * 16
    y = _run(model, x, t, plan)
    shards = torch.cat([_run(model, x[i:i + 4], t[:4], plan) for i in range(0, 16, 4)])
    assert torch.equal(shards, y)
    assert R.rel_l2(y[2:3], want[2:3]) < BF16_STEP_TOL
    d = DDPM(model=model)
    d.eval()
    steps = [0, 500, 999]
    got = d.sample((2, 8, 64, 64), seed=6, num_steps=3, schedule=steps, x_T=x[:2], progress=False).cpu()
    ref = R.ddim_sample(sd, cfg, x[:2], steps, False, 6)
    err = R.rel_l2(got, ref)
    print(f"configs[3] wide UNet 3-step DDIM vs oracle: rel-L2 {err:.3e}")
    assert err < 3e-2, err        # three bf16 steps, the first one through the 1/sqrt(abar_999) = 157x amplification
    assert_no_fault(model)


def test_config4_vae_roundtrip_512_vs_oracle():
    """configs[4]: Encoder -> Decoder on 512x512 images (latents 8x64x64): every conv level at 4x the pixel count of
    configs[1], the 64-channel halo-patch kernel on 512-wide rows.  One image against the oracle (latent and
    reconstruction), a micro-batch against its per-image runs."""
    ec, dc = R.EncoderCfg(), R.DecoderCfg()
    es, ds = R.make_encoder_state(ec, 1234), R.make_decoder_state(dc, 1234)
    enc, dec = build_encoder(ec, es, "bf16"), build_decoder(dc, ds, "bf16")
    g = torch.Generator().manual_seed(8)
    img = torch.randn(4, 3, 512, 512, generator=g).clamp(-1, 1)
    with torch.no_grad():
        z = enc(img.cuda())
        out = dec(z).cpu()
        z = z.cpu()
        z1 = torch.cat([enc(img[i:i + 1].cuda()).cpu() for i in range(4)])
        out1 = torch.cat([dec(z[i:i + 1].cuda()).cpu() for i in range(4)])
    assert tuple(z.shape) == (4, 8, 64, 64) and tuple(out.shape) == (4, 3, 512, 512)
    assert torch.equal(z1, z) and torch.equal(out1, out)
    zw = R.encoder_forward(es, ec, img[3:4])
    ow = R.decoder_forward(ds, dc, zw)
    ez = R.rel_l2(z[3:4], zw)
    eo, db = R.rel_l2(out[3:4], ow), R.psnr(out[3:4].clamp(-1, 1), ow.clamp(-1, 1))
    print(f"configs[4] 512x512 round trip: latent rel-L2 {ez:.3e}, image rel-L2 {eo:.3e}, PSNR {db:.1f} dB")
    assert ez < 1.5e-2 and eo < 2e-2 and db >= PSNR_MIN_DB
    assert_no_fault(enc); assert_no_fault(dec)
