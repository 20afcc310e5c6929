"""Helpers shared by the -m gpu parity tests (the oracle is used here only as the checker)."""
import os

import torch

from oracle import restate as R

G = os.path.join(os.path.dirname(__file__), "golden")
BF16_STEP_TOL = 1e-2     # north_star: one UNet step, relative L2, bf16 mode
FP32_STEP_TOL = 1e-5     # north_star: fp32 validation mode
PSNR_MIN_DB = 40.0       # north_star: decoded images vs the reference


def golden(name):
    return torch.load(os.path.join(G, name + ".pt"))


def build_unet(cfg: R.UNetCfg, sd, precision: str):
    from ldm_image_generator_b200 import UNet
    m = UNet(cfg.input_channels, list(cfg.stages), list(cfg.channels), cfg.stem_size)
    m.load_state_dict(sd, strict=True)
    return m.cuda().set_precision(precision)


def build_decoder(cfg: R.DecoderCfg, sd, precision: str):
    from ldm_image_generator_b200 import Decoder
    m = Decoder(cfg.output_channels, cfg.latent_channels, list(cfg.channels), list(cfg.stages))
    m.load_state_dict(sd, strict=True)
    return m.cuda().set_precision(precision)


def build_encoder(cfg: R.EncoderCfg, sd, precision: str):
    from ldm_image_generator_b200 import Encoder
    m = Encoder(cfg.input_channels, cfg.latent_channels, list(cfg.channels), list(cfg.stages))
    m.load_state_dict(sd, strict=True)
    return m.cuda().set_precision(precision)


def assert_no_fault(module):
    h = module._handle
    assert h is not None, "native library was never used"
    assert h.device_fault() == 0, "tcgen05 pipeline watchdog fired"
    assert h.launches > 0
