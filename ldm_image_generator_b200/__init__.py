"""ldm_image_generator_b200 -- the latent-diffusion sampling hot path of uthree/ldm-image-generator
(UNet step, DDIM update, VAE decode/encode) as hand-written sm_100a CUDA behind the reference's module API.

    from ldm_image_generator_b200 import DDPM, UNet, Decoder, Encoder, VAE

The arithmetic lives in libldmb200.so (csrc/, C ABI in include/ldmb.h); PyTorch is plumbing.
"""
from .ddpm import DDPM
from .unet import UNet
from .vae import VAE, Decoder, Encoder

__all__ = ["DDPM", "UNet", "Decoder", "Encoder", "VAE"]
