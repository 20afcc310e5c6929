"""Generate tests/golden/*.pt from the UNMODIFIED reference imported from /root/reference.

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference does not
travel to the GPU box):

    python oracle/gen_golden.py

Weights are the deterministic synthetic ones of ``oracle.restate.make_*_state``
(seeded torch CPU generator), loaded *strictly* into the reference modules; the
fixtures hold only the small inputs/outputs plus the Python-RNG plans, so the
GPU box can rebuild the same weights and compare against the reference's output
without the reference being present.
"""
from __future__ import annotations

import hashlib
import json
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import restate as R  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# name -> (UNetCfg kwargs, weight seed, B, H, W)
UNET_CASES = {
    "unet_tiny": (dict(input_channels=8, stages=(1, 2, 2), channels=(32, 64, 128)), 7, 2, 24, 12),
    "unet_pixel3": (dict(input_channels=3, stages=(1, 1, 2), channels=(32, 64, 64)), 8, 2, 16, 16),
    "unet_mid": (dict(input_channels=8, stages=(2, 2, 3), channels=(128, 256, 512)), 9, 2, 32, 32),
    "unet_default": (dict(), 1234, 1, 32, 32),
}
DECODER_CASES = {
    "decoder_tiny": (dict(channels=(64, 32, 16, 8)), 5, 2, 6, 10),
    "decoder_default": (dict(), 1234, 1, 8, 8),
}
ENCODER_CASES = {
    "encoder_tiny": (dict(channels=(8, 16, 32, 64)), 5, 2, 48, 32),
    "encoder_default": (dict(), 1234, 1, 64, 64),
}


def sha16(arr: np.ndarray) -> str:
    return hashlib.sha256(arr.tobytes()).hexdigest()[:16]


def main() -> None:
    sys.path.insert(0, REF)
    import unet as runet
    import vae as rvae
    import ddpm as rddpm
    os.makedirs(OUT, exist_ok=True)
    meta = {"torch": torch.__version__}

    # ---- schedule / timestep known answers straight from the reference object (ddpm.py:16-37, 66-73)
    d = rddpm.DDPM(model=runet.UNet(input_channels=8, stages=[1], channels=[32]))
    alpha = torch.cumprod(1 - d.beta, dim=0)
    sched = {"beta_sha": sha16(d.beta.numpy()), "alpha_sha": sha16(alpha.numpy()),
             "alpha_hex": {str(i): float(alpha[i]).hex() for i in (0, 500, 978, 999)}, "steps": {}}
    for n in (20, 50, 1000):
        steps = list(torch.linspace(0, d.num_timesteps - 1, n).int().numpy())
        sched["steps"][str(n)] = {"sha": sha16(np.array(steps, dtype=np.int32)),
                                  "head": [int(s) for s in steps[:8]], "tail": [int(s) for s in steps[-3:]]}
    torch.save({"beta": d.beta.clone(), "alpha": alpha.clone()}, os.path.join(OUT, "schedule.pt"))
    meta["schedule"] = sched

    # ---- Python-RNG plans as the reference consumes them (unet.py:39, modules.py:35)
    plans = {}
    probe = runet.UNet()          # default: 36 blocks
    record = []
    orig_sample, orig_random = random.sample, random.random
    for seed in (0, 1, 12345):
        for training in (False, True):
            probe.train(training)
            random.seed(seed)
            ref_plan = []
            for blk in [b for st in list(probe.encoder_stages) + list(probe.decoder_stages) for b in st.stage.blocks]:
                if blk.training and random.random() <= blk.stochastic_depth:
                    ref_plan.append([1, 0, 0]); continue
                experts = list(blk.ffn.experts)
                e1, e2 = random.sample(experts, 2)
                ref_plan.append([0, experts.index(e1), experts.index(e2)])
            plans[f"seed{seed}_{'train' if training else 'eval'}"] = {"plan": ref_plan, "next_random": random.random()}
    del probe, record, orig_sample, orig_random
    meta["plans"] = plans

    # ---- UNet forward known answers
    for name, (kw, wseed, B, H, W) in UNET_CASES.items():
        cfg = R.UNetCfg(**kw)
        sd = R.make_unet_state(cfg, wseed)
        m = runet.UNet(cfg.input_channels, list(cfg.stages), list(cfg.channels), cfg.stem_size)
        m.load_state_dict(sd, strict=True)
        g = torch.Generator().manual_seed(100 + wseed)
        x = torch.randn(B, cfg.input_channels, H, W, generator=g)
        t = torch.tensor([500, 37, 999, 0][:B])
        fix = {"x": x, "t": t}
        for training in (False, True):
            m.train(training)
            random.seed(3)
            with torch.no_grad():
                y = m(x=x, time=t, condition=None)
            random.seed(3)
            fix["plan_train" if training else "plan_eval"] = torch.tensor(R.draw_plan(len(R.block_table(cfg)), training))
            fix["y_train" if training else "y_eval"] = y.clone()
        torch.save(fix, os.path.join(OUT, name + ".pt"))
        meta[name] = dict(cfg={k: list(v) if isinstance(v, tuple) else v for k, v in kw.items()}, weight_seed=wseed,
                          y_eval_sha=sha16(fix["y_eval"].numpy()))
        print(name, "ok", tuple(fix["y_eval"].shape), float(fix["y_eval"].std()))
        del m, sd

    # ---- DDIM sampling known answer (tiny UNet, both modes, eta=0, CPU x_T from torch.manual_seed like ddpm.py:60,64)
    kw, wseed, B, H, W = UNET_CASES["unet_tiny"]
    cfg = R.UNetCfg(**kw)
    sd = R.make_unet_state(cfg, wseed)
    m = runet.UNet(cfg.input_channels, list(cfg.stages), list(cfg.channels), cfg.stem_size)
    m.load_state_dict(sd, strict=True)
    dd = rddpm.DDPM(model=m)
    fix = {}
    for training in (False, True):
        dd.train(training)
        out = dd.sample((B, cfg.input_channels, H, W), seed=11, num_steps=8, use_autocast=False)
        fix["x0_train" if training else "x0_eval"] = out.clone()
    dd.eval()
    out = dd.sample((B, cfg.input_channels, H, W), seed=11, num_steps=6, use_autocast=False,
                    schedule=[0, 40, 80, 120, 160, 199])
    fix["x0_eval_lowt"] = out.clone()
    torch.manual_seed(11)
    fix["x_T"] = torch.randn(B, cfg.input_channels, H, W)
    torch.save(fix, os.path.join(OUT, "ddim_tiny.pt"))
    print("ddim_tiny ok", float(fix["x0_eval"].std()), float(fix["x0_eval_lowt"].std()))

    # ---- VAE known answers
    for name, (kw, wseed, B, H, W) in DECODER_CASES.items():
        cfg = R.DecoderCfg(**kw)
        sd = R.make_decoder_state(cfg, wseed)
        m = rvae.Decoder(channels=list(cfg.channels)); m.load_state_dict(sd, strict=True)
        g = torch.Generator().manual_seed(200 + wseed)
        z = torch.randn(B, cfg.latent_channels, H, W, generator=g)
        with torch.no_grad():
            y = m(z)
        torch.save({"z": z, "y": y.clone()}, os.path.join(OUT, name + ".pt"))
        print(name, "ok", tuple(y.shape), float(y.std()))
    for name, (kw, wseed, B, H, W) in ENCODER_CASES.items():
        cfg = R.EncoderCfg(**kw)
        sd = R.make_encoder_state(cfg, wseed)
        m = rvae.Encoder(channels=list(cfg.channels)); m.load_state_dict(sd, strict=True)
        g = torch.Generator().manual_seed(300 + wseed)
        x = torch.randn(B, cfg.input_channels, H, W, generator=g).clamp(-1, 1)
        with torch.no_grad():
            y = m(x)
        torch.save({"x": x, "y": y.clone()}, os.path.join(OUT, name + ".pt"))
        print(name, "ok", tuple(y.shape), float(y.std()))

    with open(os.path.join(OUT, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
