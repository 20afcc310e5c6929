"""CPU: the oracle restatement reproduces the fixtures the UNMODIFIED reference produced
(oracle/gen_golden.py).  This is the pin that travels to machines without /root/reference."""
import hashlib
import json
import os
import random

import numpy as np
import pytest
import torch

from oracle import restate as R
from oracle.gen_golden import DECODER_CASES, ENCODER_CASES, UNET_CASES

G = os.path.join(os.path.dirname(__file__), "golden")
META = json.load(open(os.path.join(G, "meta.json")))
FP32_TOL = 2e-6  # fp32 CPU re-association noise between two orderings of the same maths


def sha16(a):
    return hashlib.sha256(a.tobytes()).hexdigest()[:16]


def test_schedule_bit_exact():
    """ddpm.py:19,73 tables and ddpm.py:67 timestep lists, bit for bit (SURVEY 8c hashes)."""
    beta = R.beta_table(); alpha = R.alpha_cumprod(beta)
    s = META["schedule"]
    assert sha16(beta.numpy()) == s["beta_sha"] == "a455de8584c2913e"
    assert sha16(alpha.numpy()) == s["alpha_sha"] == "b5555536933367c4"
    for i, hx in s["alpha_hex"].items():
        assert float(alpha[int(i)]).hex() == hx
    fix = torch.load(os.path.join(G, "schedule.pt"))
    assert torch.equal(fix["beta"], beta) and torch.equal(fix["alpha"], alpha)
    for n, want in (("20", "6351d40a0881c8a7"), ("50", "a2b290bba9db534e"), ("1000", "550625f47dc1b7d1")):
        steps = R.linear_steps(1000, int(n))
        assert sha16(np.array(steps, dtype=np.int32)) == s["steps"][n]["sha"] == want
    pairs = R.step_pairs(R.linear_steps(1000, 50))
    assert pairs[0] == (999, 978) and pairs[-1] == (0, 0) and pairs[-2] == (20, 0)


def test_first_ddim_iteration_scalars():
    """t=999 -> 978, eta=0."""
    co = R.ddim_coefficients(R.alpha_cumprod(R.beta_table()), 999, 978, 0.0)
    assert abs(float(co["c_div"]) - 0.006352818571) < 1e-9
    assert abs(float(co["c_eps_in"]) - 0.999979794) < 1e-7
    assert abs(float(co["c_x0"]) - 0.007837289013) < 1e-9
    assert float(co["sigma"]) == 0.0


@pytest.mark.parametrize("key", sorted(META["plans"].keys()))
def test_python_rng_plan_lockstep(key):
    """draw_plan consumes Python's random stream exactly as the reference blocks do."""
    seed = int(key.split("_")[0][4:]); training = key.endswith("train")
    random.seed(seed)
    plan = R.draw_plan(36, training)
    assert [list(p) for p in plan] == META["plans"][key]["plan"]
    assert random.random() == META["plans"][key]["next_random"]


def test_plan_known_answers():
    """random.seed(0)."""
    random.seed(0)
    assert [p[1:] for p in R.draw_plan(36, False)[:6]] == [(3, 1), (0, 1), (3, 1), (2, 1), (2, 3), (1, 2)]
    random.seed(0)
    pl = R.draw_plan(36, True)
    assert sum(p[0] for p in pl) == 10
    assert [("skip" if p[0] else p[1:]) for p in pl[:5]] == [(3, 0), (3, 1), (2, 1), (1, 2), "skip"]


def test_block_table_structure():
    """36 blocks, 8 attention blocks at decoder_stages.{0,2,3}.blocks.{1,2}, .1.blocks.{7,8}, shifts (0,3)."""
    bt = R.block_table(R.UNetCfg())
    assert len(bt) == 36
    att = [(b.prefix, b.shift) for b in bt if b.attention]
    want = [(f"decoder_stages.{i}.stage.blocks.{b}.", s) for i, bs in ((0, (1, 2)), (1, (7, 8)), (2, (1, 2)), (3, (1, 2)))
            for b, s in zip(bs, (0, 3))]
    assert att == want


@pytest.mark.parametrize("name", sorted(UNET_CASES.keys()))
def test_unet_forward_matches_reference_fixture(name):
    kw, wseed, B, H, W = UNET_CASES[name]
    cfg = R.UNetCfg(**kw)
    sd = R.make_unet_state(cfg, wseed)
    fix = torch.load(os.path.join(G, name + ".pt"))
    modes = ("eval",) if name == "unet_default" else ("eval", "train")
    for mode in modes:
        plan = [tuple(int(v) for v in row) for row in fix["plan_" + mode]]
        y = R.unet_forward(sd, cfg, fix["x"], fix["t"], plan)
        assert R.rel_l2(y, fix["y_" + mode]) < FP32_TOL, (name, mode)


def test_ddim_sampler_matches_reference_fixture():
    kw, wseed, B, H, W = UNET_CASES["unet_tiny"]
    cfg = R.UNetCfg(**kw)
    sd = R.make_unet_state(cfg, wseed)
    fix = torch.load(os.path.join(G, "ddim_tiny.pt"))
    for mode, training in (("eval", False), ("train", True)):
        x0 = R.ddim_sample(sd, cfg, fix["x_T"], R.linear_steps(1000, 8), training, py_seed=11)
        assert R.rel_l2(x0, fix["x0_" + mode]) < 5e-5, mode   # 8 steps through a 157x amplification
    x0 = R.ddim_sample(sd, cfg, fix["x_T"], [0, 40, 80, 120, 160, 199], False, py_seed=11)
    assert R.rel_l2(x0, fix["x0_eval_lowt"]) < 5e-6


@pytest.mark.parametrize("name", sorted(DECODER_CASES.keys()))
def test_decoder_matches_reference_fixture(name):
    kw, wseed, *_ = DECODER_CASES[name]
    cfg = R.DecoderCfg(**kw)
    fix = torch.load(os.path.join(G, name + ".pt"))
    y = R.decoder_forward(R.make_decoder_state(cfg, wseed), cfg, fix["z"])
    assert R.rel_l2(y, fix["y"]) < FP32_TOL


@pytest.mark.parametrize("name", sorted(ENCODER_CASES.keys()))
def test_encoder_matches_reference_fixture(name):
    kw, wseed, *_ = ENCODER_CASES[name]
    cfg = R.EncoderCfg(**kw)
    fix = torch.load(os.path.join(G, name + ".pt"))
    y = R.encoder_forward(R.make_encoder_state(cfg, wseed), cfg, fix["x"])
    assert R.rel_l2(y, fix["y"]) < FP32_TOL


def test_bilinear_and_uint8_restatements():
    import torch.nn.functional as F
    x = torch.randn(2, 3, 5, 7)
    assert torch.allclose(R.bilinear_up2(x), F.interpolate(x, scale_factor=2, mode="bilinear"), atol=1e-6)
    img = torch.tensor([[[[-2.0, -1.0, 0.0, 0.999, 1.0, 3.0]]]]).expand(1, 3, 1, 6)
    assert R.to_uint8_image(img)[0, 0, :, 0].tolist() == [0, 0, 127, 254, 255, 255]
