"""DDPM/DDIM sampler with the reference API (/root/reference/ddpm.py:10-93).

``DDPM.sample`` keeps the reference's signature, seeding, timestep list, fp32 CPU noise
schedule and torch-generator consumption; each iteration is ONE library call
(``UNet.denoise_step`` -> ``ldmb_unet_forward`` with the DDIM scalars), the posterior update
being fused into the UNet's last kernel.
"""
from __future__ import annotations

import random

import numpy as np
import torch
import torch.nn as nn
from tqdm import tqdm

from . import _lib
from .unet import UNet

_shared_default_unet = None


def _default_model() -> UNet:
    """The reference's default argument ``model=UNet()`` is evaluated once at import and shared by every
    DDPM() (ddpm.py:16).  Same sharing here, but built on first use instead of at import."""
    global _shared_default_unet
    if _shared_default_unet is None:
        _shared_default_unet = UNet()
    return _shared_default_unet


def _skip_randn(shape, device, count: int) -> None:
    """Leave torch's generator of ``device`` where ``count`` calls of ``torch.randn(*shape, device=device)`` leave it.
    A CUDA (Philox) generator advances by a fixed offset per call of a given size: one real draw measures it, the rest
    is one ``set_offset``.  Anything else (CPU generator, no offset API) just draws."""
    if count <= 0:
        return
    gen = None
    if torch.device(device).type == "cuda":
        idx = torch.device(device).index
        gen = torch.cuda.default_generators[torch.cuda.current_device() if idx is None else idx]
    done = 0
    try:
        before = gen.get_offset()
        torch.randn(*shape, device=device)
        done = 1
        step = gen.get_offset() - before
        if step > 0:
            gen.set_offset(gen.get_offset() + step * (count - 1))
            return
    except (AttributeError, RuntimeError):
        pass
    for _ in range(count - done):
        torch.randn(*shape, device=device)


class DDPM(nn.Module):
    # model(x=[B,...], time=[B], condition=None) -> eps
    def __init__(self, model=None, beta_min=1e-4, beta_max=0.02, num_timesteps=1000, loss_function=nn.L1Loss(),
                 lambda_max=20, lambda_min=-20):
        super().__init__()
        self.model = _default_model() if model is None else model
        self.num_timesteps = num_timesteps
        self.loss_function = loss_function
        self.lambda_max, self.lambda_min = lambda_max, lambda_min
        # Plain CPU fp32 tensors, not buffers: absent from state_dict and never moved by .to() (ddpm.py:19-37)
        self.beta = torch.linspace(beta_min, beta_max, num_timesteps)
        self.alpha = 1 - self.beta
        self.alpha_bar = torch.Tensor([torch.prod(self.alpha[:t]) for t in range(1, num_timesteps + 1)])
        tilde = [1]
        for t in range(1, num_timesteps):
            tilde.append((1 - self.alpha_bar[t - 1]) / (1 - self.alpha_bar[t]) * self.beta[t])
        self.beta_tilde = torch.Tensor(tilde)

    def calculate_loss(self, x, condition=None):
        raise NotImplementedError("training (ddpm.py:39-48) is outside the sampling path this package implements")

    def timesteps(self, num_steps=20, schedule="linear"):
        """ddpm.py:66-72: the descending (t, t_next) pairs the loop visits."""
        if schedule == "linear":
            steps = list(torch.linspace(0, self.num_timesteps - 1, num_steps).int().numpy())
        elif type(schedule) == list:
            steps = schedule
        else:
            # the reference does `raise f"..."`, i.e. a TypeError at run time (ddpm.py:71)
            raise TypeError(f"schedule \"{schedule}\" is not implemented.")
        return list(zip(reversed(steps), reversed([0] + steps[:-1])))

    def ddim_scalars(self, alpha_cum, t, t_next, eta):
        """The per-step fp32 scalars, computed on the CPU exactly as ddpm.py:81-85 does."""
        sigma = eta * torch.sqrt((1 - alpha_cum[t_next]) / (1 - alpha_cum[t])) * torch.sqrt(1 - alpha_cum[t] / alpha_cum[t_next])
        co = _lib.DdimCoef()
        co.c_eps_in = float(torch.sqrt(1 - alpha_cum[t]))
        co.c_div = float(torch.sqrt(alpha_cum[t]))
        co.c_x0 = float(torch.sqrt(alpha_cum[t_next]))
        co.c_eps_out = float(torch.sqrt(1 - alpha_cum[t_next] - sigma ** 2))
        co.sigma = float(sigma)
        co.final_step = int(t == 0)
        return co, sigma

    # sample as DDIM (http://arxiv.org/abs/2010.02502)
    @torch.no_grad()
    def sample(self, x_shape=(1, 3, 64, 64), condition=None, seed=None, num_steps=20, use_autocast=True,
               schedule="linear", eta=0, x_T=None, progress=True):
        """Same contract as ddpm.py:52-93.  Extra keyword-only-in-spirit arguments:
        ``x_T`` -- start from this tensor instead of ``torch.randn`` (CPU- vs GPU-generator parity,
        SURVEY.md 8c rule 4); ``progress`` -- show the tqdm bar.
        ``use_autocast`` is accepted for compatibility: arithmetic precision is the UNet's
        ``precision`` ('bf16' default, 'fp32' validation), not an autocast context."""
        device = next(self.model.parameters()).device
        if seed != None:  # noqa: E711  (reference semantics: seed=0 seeds)
            random.seed(seed)
            torch.manual_seed(seed)
            torch.cuda.manual_seed(seed)
        x = torch.randn(*x_shape, device=device)
        if x_T is not None:
            x = x_T.to(device=device, dtype=torch.float32).clone()
        pairs = self.timesteps(num_steps, schedule)
        alpha_cum = torch.cumprod((1 - self.beta), dim=0)
        fused = isinstance(self.model, UNet)
        bar = tqdm(total=len(pairs), disable=not progress)
        first = True
        chunk = self.model.film_chunk(x.shape[2] // self.model.stem_size, x.shape[3] // self.model.stem_size) if fused else 0
        # ddpm.py:80 draws the noise every step, even for eta = 0 where it is multiplied by zero: the generator must end
        # where the reference leaves it (un-reseeded multi-image loops), but nothing else consumes it inside the loop, so
        # for eta = 0 the draws are consumed up front -- one real draw, then the Philox offset is advanced
        draw_noise = not (fused and eta == 0)
        if not draw_noise:
            _skip_randn(x_shape, device, len(pairs))
        for i, (t, t_next) in enumerate(pairs):
            t, t_next = int(t), int(t_next)
            if fused and i % chunk == 0:
                # the batch-invariant Encodings MLP of the next `chunk` steps in one batched pass
                self.model.precompute_film(x, [int(p[0]) for p in pairs[i:i + chunk]])
            co, sigma = self.ddim_scalars(alpha_cum, t, t_next, eta)
            if fused:
                e = torch.randn(*x_shape, device=device) if draw_noise else None     # ddpm.py:78 then :80
                x = self.model.denoise_step(x, t, co, e if co.sigma != 0.0 else None, check_params=first)
                first = False
            else:
                # foreign eps-model: generic host-side update, same operation order as ddpm.py:82-91
                e_theta = self.model(x=x, time=torch.full((x_shape[0],), t, device=device), condition=None)
                e = torch.randn(*x_shape, device=device)
                x_t0 = (x - co.c_eps_in * e_theta) / co.c_div
                x = x_t0 if t == 0 else co.c_x0 * x_t0 + co.c_eps_out * e_theta + co.sigma * e
            if progress:
                bar.set_description(f"t: {t}, sigma: {sigma}")
            bar.update(1)
        bar.close()
        if fused and self.model._handle is not None:
            self.model._handle.raise_on_fault()     # watchdog word, polled without a stall: never hand back garbage silently
        return x

    @torch.no_grad()
    def sample_independent(self, num_images, x_shape=(1, 3, 64, 64), seeds=None, num_steps=20, use_autocast=True,
                           schedule="linear", eta=0, progress=False):
        """``[self.sample(x_shape, seed=seeds[i] if seeds else None, ...) for i in range(num_images)]`` -- the loop of
        the reference's sample scripts (sample_ldm.py:71-72 with seeds=None, sample_ddpm.py:35-36 with seeds=range(n);
        ``x_shape[0]`` must be 1 as there) -- computed as ONE batch: Python's ``random`` and torch's generator are
        consumed image by image in the order those calls would consume them (x_T, then per step the block decisions
        and the noise draw), then all images are denoised together with per-image plans (ldmb_unet_forward_per_image).
        Returns ``[num_images, *x_shape[1:]]``."""
        if x_shape[0] != 1:
            raise ValueError("sample_independent emulates batch-1 calls: x_shape[0] must be 1")
        if not isinstance(self.model, UNet):
            raise TypeError("sample_independent needs the fused UNet")
        device = next(self.model.parameters()).device
        pairs = self.timesteps(num_steps, schedule)
        n_steps = len(pairs)
        x_T, plans, noises = [], [], []
        if seeds is None:           # one uninterrupted Python-RNG stream: image-major, exactly the loop's order
            plans = self.model.draw_plans(num_images * n_steps).reshape(num_images, n_steps, -1, 3)
        for i in range(num_images):
            if seeds is not None:
                if seeds[i] != None:  # noqa: E711
                    random.seed(seeds[i])
                    torch.manual_seed(seeds[i])
                    torch.cuda.manual_seed(seeds[i])
                plans.append(self.model.draw_plans(n_steps))     # ddpm.py:78 -> unet.py:39, modules.py:35
            x_T.append(torch.randn(*x_shape, device=device))
            if eta != 0:
                noises.append([torch.randn(*x_shape, device=device) for _ in pairs])
            else:                   # ddpm.py:80 draws the noise every step even when sigma == 0: consume, do not keep
                _skip_randn(x_shape, device, n_steps)
        plans = np.asarray(plans, dtype=np.int32).reshape(num_images, n_steps, -1, 3)
        x = torch.cat(x_T, dim=0)
        alpha_cum = torch.cumprod((1 - self.beta), dim=0)
        chunk = self.model.film_chunk(x.shape[2] // self.model.stem_size, x.shape[3] // self.model.stem_size)
        bar = tqdm(total=len(pairs), disable=not progress)
        for k, (t, t_next) in enumerate(pairs):
            t, t_next = int(t), int(t_next)
            if k % chunk == 0:
                self.model.precompute_film(x, [int(p[0]) for p in pairs[k:k + chunk]])
            co, _ = self.ddim_scalars(alpha_cum, t, t_next, eta)
            e = torch.cat([noises[i][k] for i in range(num_images)], dim=0) if co.sigma != 0.0 else None
            x = self.model.denoise_step(x, t, co, e, check_params=(k == 0), plans_per_image=plans[:, k])
            bar.update(1)
        bar.close()
        self.model._handle.raise_on_fault()
        return x

