"""Build container only: pin oracle/restate.py against the imported, unmodified reference."""
import random

import pytest
import torch

from oracle import restate as R

pytestmark = pytest.mark.reference


@pytest.mark.parametrize("hw", [(16, 16), (24, 12), (28, 20), (8, 8)])
@pytest.mark.parametrize("training", [False, True])
def test_unet_restatement(reference_modules, hw, training):
    cfg = R.UNetCfg(input_channels=8, stages=(1, 2, 2), channels=(32, 64, 128))
    sd = R.make_unet_state(cfg, 7)
    m = reference_modules["unet"].UNet(8, list(cfg.stages), list(cfg.channels))
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    m.load_state_dict(sd, strict=True)
    m.train(training)
    x = torch.randn(2, 8, *hw); t = torch.tensor([500, 37])
    random.seed(3)
    with torch.no_grad():
        y_ref = m(x=x, time=t, condition=None)
    state_after = random.getstate()
    random.seed(3)
    plan = R.draw_plan(len(R.block_table(cfg)), training)
    assert random.getstate() == state_after          # Python RNG consumed in lock-step
    assert R.rel_l2(R.unet_forward(sd, cfg, x, t, plan), y_ref) < 2e-6


def test_default_state_dict_layout(reference_modules):
    """1376 tensors under model.*, none of them buffers."""
    cfg = R.UNetCfg()
    sd = R.make_unet_state(cfg)
    ref = reference_modules["ddpm"].DDPM().state_dict()
    assert len(ref) == 1376
    assert {k[len("model."):]: tuple(v.shape) for k, v in ref.items()} == {k: tuple(v.shape) for k, v in sd.items()}


def test_window_attention_restatement(reference_modules):
    torch.manual_seed(0)
    C = 64
    for shift in (0, 3):
        wa = reference_modules["attention"].WindowAttention(C, n_heads=2, window_size=6, shift=shift)
        with torch.no_grad():
            wa.attention.in_proj_bias.uniform_(-0.2, 0.2); wa.attention.out_proj.bias.uniform_(-0.2, 0.2)
        sd = {"x.attention." + k: v for k, v in wa.attention.state_dict().items()}
        for hw in ((8, 8), (13, 7), (6, 6), (4, 5), (12, 18)):
            x = torch.randn(2, C, *hw)
            with torch.no_grad():
                ref = wa(x)
            assert R.rel_l2(R.window_attention(sd, "x.", x, shift), ref) < 2e-6, (shift, hw)


def test_vae_restatement(reference_modules):
    vae = reference_modules["vae"]
    dc = R.DecoderCfg(channels=(64, 32, 16, 8)); dsd = R.make_decoder_state(dc, 5)
    dm = vae.Decoder(channels=list(dc.channels)); dm.load_state_dict(dsd, strict=True)
    z = torch.randn(2, 8, 6, 10)
    with torch.no_grad():
        assert R.rel_l2(R.decoder_forward(dsd, dc, z), dm(z)) < 2e-6
    ec = R.EncoderCfg(channels=(8, 16, 32, 64)); esd = R.make_encoder_state(ec, 5)
    em = vae.Encoder(channels=list(ec.channels)); em.load_state_dict(esd, strict=True)
    im = torch.randn(2, 3, 48, 32)
    with torch.no_grad():
        assert R.rel_l2(R.encoder_forward(esd, ec, im), em(im)) < 2e-6


def test_sampler_restatement(reference_modules):
    cfg = R.UNetCfg(input_channels=3, stages=(1, 1), channels=(32, 64))
    sd = R.make_unet_state(cfg, 21)
    m = reference_modules["unet"].UNet(3, list(cfg.stages), list(cfg.channels)); m.load_state_dict(sd)
    d = reference_modules["ddpm"].DDPM(model=m)
    for training in (False, True):
        d.train(training)
        ref = d.sample((2, 3, 8, 8), seed=5, num_steps=5, use_autocast=False)
        torch.manual_seed(5); x_T = torch.randn(2, 3, 8, 8)
        got = R.ddim_sample(sd, cfg, x_T, R.linear_steps(1000, 5), training, py_seed=5)
        assert R.rel_l2(got, ref) < 5e-5
