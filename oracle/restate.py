"""Plain-torch CPU restatement of the reference sampling path (TEST INFRASTRUCTURE ONLY).

Everything here is functional: a model is a flat ``dict`` in the reference's own
``state_dict`` key layout plus a small config object.  Each function cites the
reference lines it restates (paths relative to /root/reference).

Pinning: tests/test_oracle_vs_reference.py compares every function here with
the imported reference (where /root/reference exists); tests/test_oracle_golden.py
compares it with the fixtures committed in tests/golden/ (made by
oracle/gen_golden.py from the imported reference).  The reference has no tests
or golden vectors of its own, so those two are the pin.
"""
from __future__ import annotations

import math
import random as _pyrandom
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
State = Dict[str, Tensor]

# Constants fixed in reference code (not constructor arguments).
HEAD_DIM = 32            # unet.py:26
WINDOW = 6               # unet.py:26,51
STOCHASTIC_DEPTH = 0.25  # unet.py:26
NUM_EXPERTS = 4          # modules.py:29
NORM_EPS = 1e-4          # modules.py:19
MAX_TIMESTEPS = 10000    # sinusoidal.py:24
LEAKY_SLOPE = 0.01       # vae.py:63 (F.leaky_relu default)


# --------------------------------------------------------------------------- configs
@dataclass
class UNetCfg:
    """unet.py:75 constructor arguments."""
    input_channels: int = 8
    stages: Sequence[int] = (3, 3, 9, 3)
    channels: Sequence[int] = (128, 256, 512, 1024)
    stem_size: int = 1


@dataclass
class DecoderCfg:
    """vae.py:110 constructor arguments."""
    output_channels: int = 3
    latent_channels: int = 8
    channels: Sequence[int] = (512, 256, 128, 64)
    stages: Sequence[int] = (2, 2, 2, 2)


@dataclass
class EncoderCfg:
    """vae.py:77 constructor arguments."""
    input_channels: int = 3
    latent_channels: int = 8
    channels: Sequence[int] = (64, 128, 256, 512)
    stages: Sequence[int] = (2, 2, 2, 2)


@dataclass
class BlockInfo:
    prefix: str      # state_dict prefix, e.g. 'encoder_stages.0.stage.blocks.1.'
    level: int       # resolution level (0 = full latent resolution)
    channels: int
    attention: bool
    shift: int


def block_table(cfg: UNetCfg) -> List[BlockInfo]:
    """Blocks in execution order (unet.py:92-101, 52-61, 79-87).

    Encoder levels 0..S-1, then decoder levels S-1..0.  ``decoder_stages`` is
    stored deepest-first (insert(0, ..), unet.py:87), so list index i holds
    level S-1-i.  Block i of a stack has shift = ws//2 if i is even else 0
    (unet.py:55); only the last two blocks of a *decoder* stack attend
    (unet.py:57-60, 82, 84).
    """
    S = len(cfg.stages)
    out: List[BlockInfo] = []
    for lvl in range(S):
        for b in range(cfg.stages[lvl]):
            out.append(BlockInfo(f"encoder_stages.{lvl}.stage.blocks.{b}.", lvl, cfg.channels[lvl],
                                 False, WINDOW // 2 if b % 2 == 0 else 0))
    for i in range(S):
        lvl = S - 1 - i
        n = cfg.stages[lvl]
        for b in range(n):
            out.append(BlockInfo(f"decoder_stages.{i}.stage.blocks.{b}.", lvl, cfg.channels[lvl],
                                 b >= n - 2, WINDOW // 2 if b % 2 == 0 else 0))
    return out


# --------------------------------------------------------------------------- Python-RNG plan
def draw_plan(n_blocks: int, training: bool, rng=_pyrandom) -> List[Tuple[int, int, int]]:
    """Consume Python's ``random`` stream exactly like one UNet.forward does.

    Per block, in execution order: ``random.random()`` iff training (unet.py:39),
    then, unless skipped, ``random.sample(<4 experts>, 2)`` (modules.py:35).
    Returns (skip, e1, e2) per block.
    """
    plan = []
    for _ in range(n_blocks):
        if training and rng.random() <= STOCHASTIC_DEPTH:
            plan.append((1, 0, 0))
            continue
        e1, e2 = rng.sample(range(NUM_EXPERTS), 2)
        plan.append((0, e1, e2))
    return plan


# --------------------------------------------------------------------------- schedule (ddpm.py)
def beta_table(beta_min=1e-4, beta_max=0.02, num_timesteps=1000) -> Tensor:
    """ddpm.py:19."""
    return torch.linspace(beta_min, beta_max, num_timesteps)


def alpha_cumprod(beta: Tensor) -> Tensor:
    """ddpm.py:73 -- the table ``sample`` actually indexes (fp32 cumprod on CPU)."""
    return torch.cumprod((1 - beta), dim=0)


def linear_steps(num_timesteps: int, num_steps: int) -> List[int]:
    """ddpm.py:67: linspace(0, T-1, n).int()."""
    return [int(v) for v in torch.linspace(0, num_timesteps - 1, num_steps).int().numpy()]


def step_pairs(steps: Sequence[int]) -> List[Tuple[int, int]]:
    """ddpm.py:72,76: (t, t_next) pairs in sampling order."""
    steps = list(steps)
    nxt = [0] + steps[:-1]
    return list(zip(reversed(steps), reversed(nxt)))


def ddim_coefficients(alpha: Tensor, t: int, t_next: int, eta: float = 0.0):
    """ddpm.py:81-85: the five fp32 scalars of one update.

    x0 = (x - c_eps_in * eps) / c_div ; x' = c_x0 * x0 + c_eps_out * eps + sigma * noise
    """
    t = int(t); t_next = int(t_next)
    sigma = eta * torch.sqrt((1 - alpha[t_next]) / (1 - alpha[t])) * torch.sqrt(1 - alpha[t] / alpha[t_next])
    return dict(c_eps_in=torch.sqrt(1 - alpha[t]), c_div=torch.sqrt(alpha[t]),
                c_x0=torch.sqrt(alpha[t_next]), c_eps_out=torch.sqrt(1 - alpha[t_next] - sigma ** 2),
                sigma=sigma, final=(t == 0))


def ddim_update(x: Tensor, eps: Tensor, noise: Optional[Tensor], co) -> Tensor:
    """ddpm.py:82-91, same operation order."""
    x_t0 = (x - co["c_eps_in"] * eps) / co["c_div"]
    if co["final"]:
        return x_t0
    out = co["c_x0"] * x_t0 + co["c_eps_out"] * eps
    if noise is not None:
        out = out + co["sigma"] * noise
    else:
        out = out + co["sigma"] * torch.zeros((), dtype=x.dtype)
    return out


# --------------------------------------------------------------------------- sinusoidal.py
def position_table(C: int, H: int, W: int, dtype=torch.float32) -> Tensor:
    """sinusoidal.py:12-19 -> [C, H, W] (batch-invariant)."""
    ev = torch.arange(H, dtype=dtype).reshape(1, 1, H, 1) / H
    eh = torch.arange(W, dtype=dtype).reshape(1, 1, 1, W) / W
    factors = 1 / (2 ** (torch.arange(C // 4).reshape(1, C // 4, 1, 1) / (C // 4)))
    ev = torch.cat([torch.sin(ev * math.pi * factors), torch.cos(ev * math.pi * factors)], dim=1)
    eh = torch.cat([torch.sin(eh * math.pi * factors), torch.cos(eh * math.pi * factors)], dim=1)
    emb = torch.cat([torch.repeat_interleave(ev, W, dim=3), torch.repeat_interleave(eh, H, dim=2)], dim=1)
    return emb[0].to(dtype)


def time_table(C: int, t: Tensor, dtype=torch.float32) -> Tensor:
    """sinusoidal.py:31-39 -> [len(t), C].  t is the raw integer timestep."""
    emb = t.unsqueeze(1).expand(t.shape[0], C)
    e1, e2 = torch.chunk(emb, 2, dim=1)
    factors = 1 / (MAX_TIMESTEPS ** (torch.arange(C // 2) / (C // 2)))
    factors = factors.unsqueeze(0)
    e1 = torch.sin(e1 * math.pi * factors)
    e2 = torch.cos(e2 * math.pi * factors)
    return torch.cat([e1, e2], dim=1).to(dtype)


# --------------------------------------------------------------------------- UNet pieces
def channel_norm(x: Tensor) -> Tensor:
    """modules.py:23-25: per pixel over C, unbiased variance, eps inside the sqrt."""
    mean = x.mean(dim=1, keepdim=True)
    var = ((x - mean) ** 2).sum(dim=1, keepdim=True) / (x.shape[1] - 1)
    return (x - mean) / torch.sqrt(var + NORM_EPS)


def film_table(sd: State, p: str, C: int, H: int, W: int, t_unique: Tensor, dtype) -> Tuple[Tensor, Tensor]:
    """unet.py:18-21 evaluated once per distinct timestep (it only depends on (t,h,w)).

    Returns (mul, bias), each [len(t_unique), C, H, W].
    """
    pe = position_table(C, H, W, dtype)                         # [C,H,W]
    te = time_table(C, t_unique, dtype)                         # [T,C]
    T = t_unique.shape[0]
    embs = torch.cat([pe.unsqueeze(0).expand(T, C, H, W), te[:, :, None, None].expand(T, C, H, W)], dim=1)
    h = F.relu(F.conv2d(embs, sd[p + "encodings.proj1.weight"], sd[p + "encodings.proj1.bias"]))
    h = F.conv2d(h, sd[p + "encodings.proj2.weight"], sd[p + "encodings.proj2.bias"])
    mul, bias = torch.chunk(h, 2, dim=1)
    return mul, bias


def reglu(sd: State, p: str, x: Tensor) -> Tensor:
    """modules.py:14-15: c(a(x) * relu(b(x))), all 1x1 convs."""
    a = F.conv2d(x, sd[p + "a.weight"], sd[p + "a.bias"])
    b = F.conv2d(x, sd[p + "b.weight"], sd[p + "b.bias"])
    return F.conv2d(a * F.relu(b), sd[p + "c.weight"], sd[p + "c.bias"])


def _mha_core(tokens: Tensor, key_bias: Optional[Tensor], w_in, b_in, w_out, b_out, heads: int) -> Tensor:
    """torch F.multi_head_attention_forward, need_weights=True branch, as called
    at attention.py:82: packed in-projection, q scaled by sqrt(1/d) *before*
    QK^T, key bias added to the logits, softmax, PV, out-projection.

    tokens [N, L, C]; key_bias [N, L] (float, may hold -inf) or None.
    """
    N, L, C = tokens.shape
    d = C // heads
    qkv = tokens @ w_in.t() + b_in
    q, k, v = qkv.split(C, dim=-1)
    q = q.reshape(N, L, heads, d).transpose(1, 2) * math.sqrt(1.0 / float(d))
    k = k.reshape(N, L, heads, d).transpose(1, 2)
    v = v.reshape(N, L, heads, d).transpose(1, 2)
    logits = q @ k.transpose(-2, -1)                             # [N, heads, L, L]
    if key_bias is not None:
        logits = logits + key_bias[:, None, None, :]
    w = torch.softmax(logits, dim=-1)
    o = (w @ v).transpose(1, 2).reshape(N, L, C)
    return o @ w_out.t() + b_out


def window_attention(sd: State, p: str, x: Tensor, shift: int) -> Tensor:
    """attention.py:13-85 by index arithmetic instead of pad/roll/split copies.

    Work in the zero-padded frame xp (attention.py:27-28).  A padded position
    q=(i,j) sits, after the roll (attention.py:39), at r=((i+s)%Hp,(j+s)%Wp) and
    belongs to window (r_i//ws, r_j//ws).  Keys of a query are all padded
    positions of its window.  Key bias:
      shift == 0: -inf on pad positions (bool key_padding_mask, attention.py:33-35);
      shift != 0: attention.py:40 rolls the *already rolled activation* into
        ``mask``, so the float value added to the logits of key q is channel 0
        of xp at ((i-s)%Hp,(j-s)%Wp); pad keys are then live with q/k/v = bias.
    Outputs at pad positions are cropped (attention.py:56).
    """
    C = x.shape[1]
    heads = C // HEAD_DIM
    w_in, b_in = sd[p + "attention.in_proj_weight"], sd[p + "attention.in_proj_bias"]
    w_out, b_out = sd[p + "attention.out_proj.weight"], sd[p + "attention.out_proj.bias"]
    B, _, H, W = x.shape
    ws = WINDOW
    if H <= ws and W <= ws:                                       # attention.py:15-16
        tok = x.flatten(2).transpose(1, 2)
        return _mha_core(tok, None, w_in, b_in, w_out, b_out, heads).transpose(1, 2).reshape(B, C, H, W)
    Hp, Wp = -(-H // ws) * ws, -(-W // ws) * ws
    xp = F.pad(x, (0, Wp - W, 0, Hp - H))
    ii, jj = torch.meshgrid(torch.arange(Hp), torch.arange(Wp), indexing="ij")
    ri, rj = (ii + shift) % Hp, (jj + shift) % Wp
    win = (ri // ws) * (Wp // ws) + (rj // ws)                    # window id of each padded position
    slot = (ri % ws) * ws + (rj % ws)
    nwin, L = (Hp // ws) * (Wp // ws), ws * ws
    order = torch.empty(nwin * L, dtype=torch.long)
    order[(win * L + slot).flatten()] = torch.arange(Hp * Wp)     # flat padded index of (window, slot)
    flat = xp.flatten(2)                                          # [B, C, Hp*Wp]
    tok = flat[:, :, order].reshape(B, C, nwin, L).permute(0, 2, 3, 1).reshape(B * nwin, L, C)
    if shift == 0:
        pad = ((ii >= H) | (jj >= W)).flatten()[order]
        kb = torch.zeros(nwin * L, dtype=x.dtype).masked_fill(pad, float("-inf"))
        kb = kb.reshape(1, nwin, L).expand(B, nwin, L).reshape(B * nwin, L)
    else:
        src = (((ii - shift) % Hp) * Wp + ((jj - shift) % Wp)).flatten()[order]
        kb = flat[:, 0, :][:, src].reshape(B * nwin, L)
    o = _mha_core(tok, kb, w_in, b_in, w_out, b_out, heads)       # [B*nwin, L, C]
    o = o.reshape(B, nwin * L, C).transpose(1, 2)                 # [B, C, nwin*L]
    out = torch.empty_like(flat)
    out[:, :, order] = o
    return out.reshape(B, C, Hp, Wp)[:, :, :H, :W]


def swin_block(sd: State, info: BlockInfo, x: Tensor, t: Tensor, plan_entry) -> Tensor:
    """unet.py:38-48 + modules.py:34-36 with the Python-RNG decisions made explicit."""
    skip, e1, e2 = plan_entry
    if skip:
        return x
    p = info.prefix
    B, C, H, W = x.shape
    t_unique, inverse = torch.unique(t, return_inverse=True)
    mul, bias = film_table(sd, p, C, H, W, t_unique, x.dtype)
    xm = channel_norm(x) * mul[inverse] + bias[inverse]
    y = (reglu(sd, p + "ffn.general.", xm) + reglu(sd, p + f"ffn.experts.{e1}.", xm)
         + reglu(sd, p + f"ffn.experts.{e2}.", xm))
    y = y + F.conv2d(xm, sd[p + "conv.weight"], sd[p + "conv.bias"], padding=1, groups=C // HEAD_DIM)
    if info.attention:
        y = y + window_attention(sd, p + "self_attention.", xm, info.shift)
    return y + x


def unet_forward(sd: State, cfg: UNetCfg, x: Tensor, t: Tensor, plan) -> Tensor:
    """unet.py:89-103.  ``plan`` = draw_plan(...) for this call."""
    S = len(cfg.stages)
    blocks = block_table(cfg)
    assert len(plan) == len(blocks)
    s = cfg.stem_size
    x = F.conv2d(x, sd["encoder_first.weight"], sd["encoder_first.bias"], stride=s)
    skips: List[object] = []
    bi = 0
    for lvl in range(S):
        for _ in range(cfg.stages[lvl]):
            x = swin_block(sd, blocks[bi], x, t, plan[bi]); bi += 1
        if lvl == S - 1:
            skips.insert(0, 0)
        else:
            skips.insert(0, x)
            x = F.conv2d(x, sd[f"encoder_stages.{lvl}.ch_conv.0.weight"], sd[f"encoder_stages.{lvl}.ch_conv.0.bias"])
            x = F.avg_pool2d(x, 2)
    for i in range(S):
        lvl = S - 1 - i
        if lvl != S - 1:
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            x = F.conv2d(x, sd[f"decoder_stages.{i}.ch_conv.1.weight"], sd[f"decoder_stages.{i}.ch_conv.1.bias"])
        x = x + skips[i]
        for _ in range(cfg.stages[lvl]):
            x = swin_block(sd, blocks[bi], x, t, plan[bi]); bi += 1
    return F.conv_transpose2d(x, sd["decoder_last.weight"], sd["decoder_last.bias"], stride=s)


# --------------------------------------------------------------------------- sampler (ddpm.py:51-93)
def ddim_sample(sd: State, cfg: UNetCfg, x_T: Tensor, steps: Sequence[int], training: bool,
                py_seed: Optional[int], eta: float = 0.0, alpha: Optional[Tensor] = None,
                noises: Optional[Sequence[Tensor]] = None, trajectory: Optional[list] = None) -> Tensor:
    """The loop of DDPM.sample with x_T given explicitly.

    ``py_seed`` seeds Python's ``random`` (ddpm.py:58) when not None.  With
    eta == 0 the per-step noise is multiplied by zero and is not needed.
    ``trajectory`` (if a list) receives (t, x_in, eps) per step.
    """
    if alpha is None:
        alpha = alpha_cumprod(beta_table())
    if py_seed is not None:
        _pyrandom.seed(py_seed)
    n_blocks = len(block_table(cfg))
    x = x_T
    for k, (t, t_next) in enumerate(step_pairs(steps)):
        plan = draw_plan(n_blocks, training)
        tt = torch.full((x.shape[0],), int(t), dtype=torch.long)
        eps = unet_forward(sd, cfg, x, tt, plan)
        if trajectory is not None:
            trajectory.append((int(t), x, eps))
        co = ddim_coefficients(alpha, t, t_next, eta)
        x = ddim_update(x, eps, None if noises is None else noises[k], co)
    return x


# --------------------------------------------------------------------------- VAE (vae.py)
def res_block(sd: State, p: str, x: Tensor) -> Tensor:
    """vae.py:60-66."""
    y = F.leaky_relu(F.conv2d(x, sd[p + "c1.weight"], sd[p + "c1.bias"], padding=1), LEAKY_SLOPE)
    y = F.leaky_relu(F.conv2d(y, sd[p + "c2.weight"], sd[p + "c2.bias"], padding=1), LEAKY_SLOPE)
    return y + x


def bilinear_up2(x: Tensor) -> Tensor:
    """F.interpolate(scale_factor=2, mode='bilinear') == align_corners=False (vae.py:131),
    written out: out[2i] = .25*x[i-1] + .75*x[i], out[2i+1] = .75*x[i] + .25*x[i+1], edges clamped,
    separably along H then W."""
    def up(v: Tensor, dim: int) -> Tensor:
        n = v.shape[dim]
        idx = torch.arange(n)
        prev = v.index_select(dim, (idx - 1).clamp(min=0))
        nxt = v.index_select(dim, (idx + 1).clamp(max=n - 1))
        even = 0.25 * prev + 0.75 * v
        odd = 0.75 * v + 0.25 * nxt
        return torch.stack([even, odd], dim=dim + 1).flatten(dim, dim + 1)
    return up(up(x, 2), 3)


def decoder_forward(sd: State, cfg: DecoderCfg, z: Tensor) -> Tensor:
    """vae.py:122-132 (output_layer is never applied, vae.py:113)."""
    x = F.conv2d(z, sd["input_layer.weight"], sd["input_layer.bias"])
    rgb_out = None
    for s, n in enumerate(cfg.stages):
        if s > 0:
            x = F.conv_transpose2d(x, sd[f"upsamples.{s}.weight"], sd[f"upsamples.{s}.bias"], stride=2)
        for l in range(n):
            x = res_block(sd, f"stages.{s}.layers.{l}.", x)
        rgb = F.conv2d(x, sd[f"stages.{s}.to_rgb.weight"], sd[f"stages.{s}.to_rgb.bias"])
        rgb_out = rgb if rgb_out is None else bilinear_up2(rgb_out) + rgb
    return rgb_out


def encoder_forward(sd: State, cfg: EncoderCfg, x: Tensor) -> Tensor:
    """vae.py:91-96."""
    x = F.conv2d(x, sd["input_layer.weight"], sd["input_layer.bias"])
    S = len(cfg.stages)
    for s, n in enumerate(cfg.stages):
        for l in range(n):
            x = res_block(sd, f"stages.{s}.seq.{l}.", x)
        if s != S - 1:
            x = F.avg_pool2d(x, 2)
            x = F.conv2d(x, sd[f"downsamples.{s}.1.weight"], sd[f"downsamples.{s}.1.bias"])
    return F.conv2d(x, sd["output_layer.weight"], sd["output_layer.bias"])


def to_uint8_image(img: Tensor) -> Tensor:
    """sample_ldm.py:75-77: clamp, *127.5+127.5, truncate to uint8, CHW->HWC (per image)."""
    img = torch.clamp(img, -1, 1)
    return (img * 127.5 + 127.5).to(torch.uint8).permute(0, 2, 3, 1).contiguous()


# --------------------------------------------------------------------------- synthetic weights
def _uniform(gen: torch.Generator, shape, bound: float) -> Tensor:
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1) * bound


def _conv(sd: State, gen, name: str, cout: int, cin: int, k: int, transposed: bool = False):
    fan_in = cin * k * k
    shape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
    sd[name + ".weight"] = _uniform(gen, shape, 1.0 / math.sqrt(fan_in))
    sd[name + ".bias"] = _uniform(gen, (cout,), 1.0 / math.sqrt(fan_in))


def make_unet_state(cfg: UNetCfg, seed: int = 1234) -> State:
    """Deterministic synthetic UNet weights in the reference key layout.

    U(+-1/sqrt(fan_in)) like torch's default conv init, but with *non-zero*
    attention biases so pad-token handling is exercised.  Same torch build =>
    same bits here and on the GPU box.
    """
    gen = torch.Generator().manual_seed(seed)
    sd: State = {}
    s = cfg.stem_size
    _conv(sd, gen, "encoder_first", cfg.channels[0], cfg.input_channels, s)
    # ConvTranspose2d(channels[0] -> input_channels): weight [in=C0, out=Cin, k, k]
    sd["decoder_last.weight"] = _uniform(gen, (cfg.channels[0], cfg.input_channels, s, s), 1.0 / math.sqrt(cfg.channels[0]))
    sd["decoder_last.bias"] = _uniform(gen, (cfg.input_channels,), 1.0 / math.sqrt(cfg.channels[0]))
    for info in block_table(cfg):
        p, C = info.prefix, info.channels
        for e in ["ffn.general"] + [f"ffn.experts.{i}" for i in range(NUM_EXPERTS)]:
            for m in "abc":
                _conv(sd, gen, f"{p}{e}.{m}", C, C, 1)
        fan = HEAD_DIM * 9
        sd[p + "conv.weight"] = _uniform(gen, (C, HEAD_DIM, 3, 3), 1.0 / math.sqrt(fan))
        sd[p + "conv.bias"] = _uniform(gen, (C,), 1.0 / math.sqrt(fan))
        _conv(sd, gen, p + "encodings.proj1", 4 * C, 2 * C, 1)
        _conv(sd, gen, p + "encodings.proj2", 2 * C, 4 * C, 1)
        if info.attention:
            for a in ("self_attention", "cross_attention"):
                q = f"{p}{a}.attention."
                sd[q + "in_proj_weight"] = _uniform(gen, (3 * C, C), math.sqrt(6.0 / (4 * C)))
                sd[q + "in_proj_bias"] = _uniform(gen, (3 * C,), 0.1)
                sd[q + "out_proj.weight"] = _uniform(gen, (C, C), 1.0 / math.sqrt(C))
                sd[q + "out_proj.bias"] = _uniform(gen, (C,), 0.1)
    S = len(cfg.stages)
    for lvl in range(S - 1):
        _conv(sd, gen, f"encoder_stages.{lvl}.ch_conv.0", cfg.channels[lvl + 1], cfg.channels[lvl], 1)
        i = S - 1 - lvl
        _conv(sd, gen, f"decoder_stages.{i}.ch_conv.1", cfg.channels[lvl], cfg.channels[lvl + 1], 1)
    return sd


def make_decoder_state(cfg: DecoderCfg, seed: int = 1234) -> State:
    gen = torch.Generator().manual_seed(seed + 1)
    sd: State = {}
    _conv(sd, gen, "input_layer", cfg.channels[0], cfg.latent_channels, 1)
    _conv(sd, gen, "output_layer", cfg.output_channels, cfg.channels[-1], 1)
    for s, (c, n) in enumerate(zip(cfg.channels, cfg.stages)):
        for l in range(n):
            _conv(sd, gen, f"stages.{s}.layers.{l}.c1", c, c, 3)
            _conv(sd, gen, f"stages.{s}.layers.{l}.c2", c, c, 3)
        _conv(sd, gen, f"stages.{s}.to_rgb", cfg.output_channels, c, 1)
        if s > 0:
            _conv(sd, gen, f"upsamples.{s}", c, cfg.channels[s - 1], 2, transposed=True)
    return sd


def make_encoder_state(cfg: EncoderCfg, seed: int = 1234) -> State:
    gen = torch.Generator().manual_seed(seed + 2)
    sd: State = {}
    _conv(sd, gen, "input_layer", cfg.channels[0], cfg.input_channels, 1)
    _conv(sd, gen, "output_layer", cfg.latent_channels, cfg.channels[-1], 1)
    for s, (c, n) in enumerate(zip(cfg.channels, cfg.stages)):
        for l in range(n):
            _conv(sd, gen, f"stages.{s}.seq.{l}.c1", c, c, 3)
            _conv(sd, gen, f"stages.{s}.seq.{l}.c2", c, c, 3)
        if s != len(cfg.stages) - 1:
            _conv(sd, gen, f"downsamples.{s}.1", cfg.channels[s + 1], c, 1)
    return sd


# --------------------------------------------------------------------------- metrics
def rel_l2(a: Tensor, b: Tensor) -> float:
    a = a.double(); b = b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def psnr(a: Tensor, b: Tensor, peak: float = 2.0) -> float:
    """PSNR over images in [-1, 1] (peak-to-peak 2)."""
    mse = float(((a.double() - b.double()) ** 2).mean())
    return float("inf") if mse == 0 else 10.0 * math.log10(peak * peak / mse)
