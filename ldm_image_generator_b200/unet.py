"""UNet with the reference constructor / forward / state_dict contract
(/root/reference/unet.py:9-103), executed by libldmb200.so.

``UNet.forward(x, time, condition=None)`` makes ONE C-ABI call (``ldmb_unet_forward``) that
runs the whole network on the current CUDA stream.  Host code here only
  * draws the Python-``random`` decisions of the step in the reference's order
    (stochastic depth unet.py:39, expert picks modules.py:35) into a plan,
  * evaluates the (t)-only and (h,w)-only sinusoid tables (sinusoidal.py) on the host,
  * keeps the library's repacked weight arena in sync with the nn.Parameters.
"""
from __future__ import annotations

import ctypes
import random
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import _lib, runtime
from .attention import CrossAttention, WindowAttention
from .modules import ChannelNorm, RandomMoE, _FusedIntoUNet
from .sinusoidal import PositionalEncoding2d, TimeEncoding2d


class Encodings(_FusedIntoUNet):
    """Time + position FiLM (unet.py:9-23).  Depends on (t, h, w) only, so the library evaluates
    the MLP once per call for all images instead of once per image."""

    def __init__(self, channels):
        super().__init__()
        self.proj1 = nn.Conv2d(channels * 2, channels * 4, 1, 1, 0)
        self.act = nn.ReLU()
        self.proj2 = nn.Conv2d(channels * 4, channels * 2, 1, 1, 0)
        self.pe = PositionalEncoding2d(channels, return_encoding_only=True)
        self.te = TimeEncoding2d(channels, return_encoding_only=True)


class SwinBlock(_FusedIntoUNet):
    def __init__(self, channels, head_dim=32, window_size=6, shift=0, attention=True, stochastic_depth=0.25):
        super().__init__()
        if head_dim != 32 or window_size != 6:
            raise runtime.LdmbError("libldmb200 implements SwinBlock as the reference instantiates it: head_dim=32, window 6")
        self.norm = ChannelNorm(channels)
        self.ffn = RandomMoE(channels)
        self.conv = nn.Conv2d(channels, channels, 3, 1, 1, groups=channels // head_dim)
        self.stochastic_depth = stochastic_depth
        self.attention_flag = attention
        if attention:
            self.self_attention = WindowAttention(channels, n_heads=channels // head_dim, window_size=window_size, shift=shift)
            self.cross_attention = CrossAttention(channels, n_heads=channels // head_dim)
        self.encodings = Encodings(channels)
        self.shift = shift


class SwinStack(_FusedIntoUNet):
    def __init__(self, channels, head_dim=32, window_size=6, num_blocks=2, attention=True):
        super().__init__()
        blocks = []
        for i in range(num_blocks):
            blocks.append(SwinBlock(channels, head_dim, window_size,
                                    shift=window_size // 2 if i % 2 == 0 else 0,
                                    attention=attention and i >= num_blocks - 2))   # unet.py:55-60
        self.blocks = nn.ModuleList(blocks)


class UNetBlock(nn.Module):
    def __init__(self, stage, ch_conv):
        super().__init__()
        self.stage = stage
        self.ch_conv = ch_conv


class UNet(nn.Module):
    def __init__(self, input_channels=8, stages=[3, 3, 9, 3], channels=[128, 256, 512, 1024], stem_size=1):
        super().__init__()
        self.input_channels, self.stem_size = input_channels, stem_size
        self.stage_blocks, self.stage_channels = list(stages), list(channels)
        self.encoder_first = nn.Conv2d(input_channels, channels[0], stem_size, stem_size, 0)
        self.decoder_last = nn.ConvTranspose2d(channels[0], input_channels, stem_size, stem_size, 0)
        self.encoder_stages = nn.ModuleList([])
        self.decoder_stages = nn.ModuleList([])
        last = len(stages) - 1
        for i, (n, c) in enumerate(zip(stages, channels)):
            # same construction order as unet.py:80-87 so a torch seed yields the reference's initial weights
            enc = SwinStack(c, num_blocks=n, attention=False)
            down = nn.Identity() if i == last else nn.Sequential(nn.Conv2d(c, channels[i + 1], 1, 1, 0), nn.AvgPool2d(kernel_size=2))
            dec = SwinStack(c, num_blocks=n)
            up = nn.Identity() if i == last else nn.Sequential(nn.Upsample(scale_factor=2), nn.Conv2d(channels[i + 1], c, 1, 1, 0))
            self.encoder_stages.append(UNetBlock(enc, down))
            self.decoder_stages.insert(0, UNetBlock(dec, up))     # deepest first (unet.py:87)
        self.precision = runtime.default_precision()
        self._handle: Optional[runtime.Handle] = None
        self._pe_res = None
        self._film = None      # (handle, (H, W), {t: row}) of the FiLM tables precomputed for a schedule
        self._deterministic = None     # None: the library default (environment variable LDMB_DETERMINISTIC)

    # ------------------------------------------------------------------ host-side helpers
    def set_precision(self, precision: str) -> "UNet":
        """'bf16' (tcgen05 path) or 'fp32' (CUDA-core validation mode)."""
        if precision not in runtime.PRECISIONS:
            raise ValueError(precision)
        if precision != self.precision:
            self.precision, self._handle, self._pe_res, self._film = precision, None, None, None
        return self

    def set_deterministic(self, on: bool = True) -> "UNet":
        """bf16 mode: bit-reproducible (and batch-size independent) results at a few % of speed -- no split-K GEMM
        slices, no concurrent grouped-conv branch, so every element of the residual stream is updated in a fixed order."""
        self._deterministic = bool(on)
        if self._handle is not None:
            self._handle.set_deterministic(self._deterministic)
        return self

    _RUNTIME_STATE = ("_handle", "_pe_res", "_film")

    def __getstate__(self):
        """copy.deepcopy / pickle / torch.save(module): the device handle (ctypes pointers) and the caches tied to it are
        per-object runtime state, rebuilt on first use -- the reference module supports all three."""
        state = self.__dict__.copy()
        for k in self._RUNTIME_STATE:
            state[k] = None
        state.pop("_param_items", None)
        state.pop("_te_cache", None)
        return state

    def invalidate_weights(self) -> None:
        """Force a re-upload of every parameter (and a recomputation of the FiLM tables) at the next call.  Parameter
        changes are detected by (storage, ``_version``, shape, dtype); writes THROUGH ``.data`` (``p.data.copy_()``, common
        in EMA / weight-averaging code) do not bump ``_version`` -- call this after them."""
        if self._handle is not None:
            self._handle._param_keys.clear()
        self._film = None
        self.__dict__.pop("_param_items", None)

    def blocks_in_execution_order(self) -> List[SwinBlock]:
        out = []
        for st in list(self.encoder_stages) + list(self.decoder_stages):
            out.extend(st.stage.blocks)
        return out

    def draw_plan(self) -> List[Sequence[int]]:
        """Consume Python's ``random`` exactly like one reference forward: per block, ``random.random()``
        iff the block is in training mode (unet.py:39), then ``random.sample`` of 2 of the 4 experts
        (modules.py:35) unless skipped."""
        plan = []
        for blk in self.blocks_in_execution_order():
            if blk.training and random.random() <= blk.stochastic_depth:
                plan.append((1, 0, 0))
                continue
            e1, e2 = random.sample(range(len(blk.ffn.experts)), 2)
            plan.append((0, e1, e2))
        return plan

    _replay_checked: Optional[bool] = None       # class-wide: has the MT19937 replay been checked against `random`?

    def draw_plans(self, n: int) -> np.ndarray:
        """``[self.draw_plan() for _ in range(n)]`` as an int32 ``[n, n_blocks, 3]`` array, with Python's ``random`` left in
        the state those calls leave it in.  The per-image sampler needs ``images x steps`` plans before its first step
        (the batch-1 loop draws image 0's whole trajectory before image 1's first step), ~0.2 s of interpreter time for
        64 x 50; here the decisions are replayed in C (``ldmb_host_draw_plans``) from the raw MT19937 stream of the
        generator's current state.  The replay is checked once per process against the ``random`` module itself
        (results and end state) and abandoned for the plain loop if this interpreter's ``random`` consumes differently."""
        cls = UNet
        if cls._replay_checked is None:
            saved = random.getstate()
            slow = np.array([self.draw_plan() for _ in range(3)], dtype=np.int32)
            end = random.getstate()
            random.setstate(saved)
            fast = self._replay_plans(3)
            cls._replay_checked = fast is not None and np.array_equal(fast, slow) and random.getstate() == end
            random.setstate(saved)
        out = self._replay_plans(n) if cls._replay_checked else None
        if out is None:
            out = np.array([self.draw_plan() for _ in range(n)], dtype=np.int32).reshape(n, -1, 3)
        return out

    def _replay_plans(self, n: int) -> Optional[np.ndarray]:
        blocks = self.blocks_in_execution_order()
        n_experts = {len(b.ffn.experts) for b in blocks}
        version, internal, gauss = random.getstate()
        if len(n_experts) != 1 or not 2 <= min(n_experts) <= 21 or version != 3 or len(internal) != 625:
            return None
        key, pos = np.array(internal[:-1], dtype=np.uint32), int(internal[-1])
        training = np.array([bool(b.training) for b in blocks], dtype=np.uint8)
        depth = np.array([float(b.stochastic_depth) for b in blocks], dtype=np.float64)
        out = np.empty((n, len(blocks), 3), dtype=np.int32)
        used = ctypes.c_int64(0)
        lib, mt = _lib.load(), np.random.MT19937()
        n_raw = 6 * n * len(blocks) + 64
        while True:
            mt.state = {"bit_generator": "MT19937", "state": {"key": key, "pos": pos}}
            raw = np.ascontiguousarray(mt.random_raw(n_raw), dtype=np.uint32)
            rc = lib.ldmb_host_draw_plans(raw.ctypes.data, n_raw, n, len(blocks), training.ctypes.data, depth.ctypes.data,
                                          min(n_experts), out.ctypes.data, ctypes.byref(used))
            if rc == 0:
                break
            if n_raw > (1 << 30):
                return None
            n_raw *= 2                                # rejection sampling ran past the stream: replay with a longer one
        mt.state = {"bit_generator": "MT19937", "state": {"key": key, "pos": pos}}
        if used.value:
            mt.random_raw(used.value)
        st = mt.state["state"]
        random.setstate((version, tuple(int(v) for v in st["key"]) + (int(st["pos"]),), gauss))
        return out

    def _prepare(self, device: torch.device, check_params: bool = True) -> runtime.Handle:
        """The device handle, with every changed parameter re-uploaded.  Walking the 1 376-entry state_dict costs
        milliseconds of host time, so DDPM.sample asks for it once per call (``check_params=False`` on the
        remaining steps: weights cannot change inside the no_grad sampling loop)."""
        h = self._handle
        if h is None or h.device != torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device()):
            h = runtime.Handle(device, self.precision)
            h.unet_configure(self.input_channels, self.stage_blocks, self.stage_channels, self.stem_size)
            if self._deterministic is not None:
                h.set_deterministic(self._deterministic)
            self._handle, self._pe_res, self._film = h, None, None
            check_params = True
        if check_params:
            items = self.__dict__.get("_param_items")
            if items is None:      # (name, tensor) pairs; dropped whenever the module tree may have new tensors
                items = self.__dict__["_param_items"] = list(self.state_dict(keep_vars=True).items())
            if h.unet_load(items):
                self._film = None      # precomputed FiLM tables are stale
        return h

    def _set_position_tables(self, h: runtime.Handle, Hs: int, Ws: int) -> None:
        if self._pe_res != (Hs, Ws):
            for lvl, c in enumerate(self.stage_channels):
                h.unet_set_position_table(lvl, self.encoder_stages[lvl].stage.blocks[0].encodings.pe.table(Hs >> lvl, Ws >> lvl))
            self._pe_res = (Hs, Ws)
            self._film = None

    FILM_BYTES_BUDGET = 3 << 30       # device memory for the per-schedule FiLM tables + their scratch

    def film_chunk(self, Hs: int, Ws: int) -> int:
        """How many timesteps of FiLM tables fit the budget (tables fp32 [blocks][t][HW][2C] + bf16 scratch of equal size)."""
        per_t = 0
        for lvl, (n, c) in enumerate(zip(self.stage_blocks, self.stage_channels)):
            per_t += 2 * n * (Hs >> lvl) * (Ws >> lvl) * 2 * c * (4 + 4)
        return max(1, self.FILM_BYTES_BUDGET // max(per_t, 1))

    def precompute_film(self, x: torch.Tensor, timesteps: Sequence[int]) -> None:
        """Evaluate the Encodings MLP (unet.py:18-21) of every block for all `timesteps` at x's resolution in one
        batched pass; denoise steps at these timesteps then skip it (it depends on (t, h, w) only)."""
        runtime._require_cuda(x, "UNet input")
        with torch.cuda.device(x.device):
            h = self._prepare(x.device, True)
            H, W = x.shape[2], x.shape[3]
            self._set_position_tables(h, H // self.stem_size, W // self.stem_size)
            uniq = sorted(set(int(t) for t in timesteps))
            h.unet_precompute_film(H, W, self._time_tables(uniq))
            self._film = (h, (H, W), {t: i for i, t in enumerate(uniq)})

    def _apply(self, fn, *args, **kwargs):          # .to() / .cuda() / .float(): tensors may be replaced
        self.__dict__.pop("_param_items", None)
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):     # assign=True replaces the Parameter objects
        self.__dict__.pop("_param_items", None)
        return super().load_state_dict(*args, **kwargs)

    def _time_tables(self, uniq: Sequence[int]):
        """Per-level sinusoidal time tables (sinusoidal.py:31-41) of the distinct timesteps, cached per tuple."""
        key = tuple(uniq)
        cache = self.__dict__.setdefault("_te_cache", {})
        te = cache.get(key)
        if te is None:
            if len(cache) >= 4096:
                cache.clear()
            tt = torch.tensor(list(uniq), dtype=torch.long)
            te = [self.encoder_stages[lvl].stage.blocks[0].encodings.te.table(tt) for lvl in range(len(self.stage_channels))]
            cache[key] = te
        return te

    def _run(self, x: torch.Tensor, t_values: Sequence[int], coef=None, noise=None, out=None, plan=None,
             check_params: bool = True, plans_per_image=None) -> torch.Tensor:
        """``plan``: one (skip, e1, e2) per block shared by the batch (one reference forward on the batch);
        ``plans_per_image``: one such plan PER IMAGE (a batch of the reference's batch-1 forwards)."""
        runtime._require_cuda(x, "UNet input")
        p0 = next(self.parameters())
        runtime._require_cuda(p0, "UNet parameters")
        if x.dim() != 4 or x.shape[1] != self.input_channels:
            raise RuntimeError(f"expected input[{', '.join(map(str, x.shape))}] to have {self.input_channels} channels, "
                               f"but got {x.shape[1] if x.dim() == 4 else '?'} channels instead")
        x = runtime.f32c(x)
        with torch.cuda.device(x.device):
            h = self._prepare(x.device, check_params)
            B, _, H, W = x.shape
            Hs, Ws = H // self.stem_size, W // self.stem_size
            self._set_position_tables(h, Hs, Ws)
            film = self._film
            if film is not None and film[0] is h and film[1] == (H, W) and all(int(v) in film[2] for v in t_values):
                index, te = film[2], len(film[2])           # FiLM tables of these timesteps are already on the device
            else:
                self._film = None                           # this call overwrites the FiLM workspace
                uniq = sorted(set(int(v) for v in t_values))
                index = {v: i for i, v in enumerate(uniq)}
                te = self._time_tables(uniq)
            if out is None:
                out = torch.empty_like(x)
            if plans_per_image is not None:
                if len(plans_per_image) != B:
                    raise RuntimeError(f"plans_per_image has {len(plans_per_image)} entries for a batch of {B}")
                by_block = np.ascontiguousarray(np.asarray(plans_per_image, dtype=np.int32).transpose(1, 0, 2))
                h.unet_forward(x, out, [index[int(v)] for v in t_values], te, by_block, coef, noise, per_image=True)
                return out
            if plan is None:
                plan = self.draw_plan()
            h.unet_forward(x, out, [index[int(v)] for v in t_values], te, plan, coef, noise)
        return out

    # ------------------------------------------------------------------ reference API
    def forward(self, x, time, condition=None):
        """eps = UNet(x, time).  ``condition`` is accepted and ignored, as in the reference
        (unet.py:93,101 never pass it to the stages)."""
        t_values = [int(v) for v in time.detach().reshape(-1).tolist()]
        if len(t_values) == 1 and x.shape[0] > 1:
            t_values = t_values * x.shape[0]
        return self._run(x, t_values)

    def forward_independent(self, x, time):
        """The batch as B separate reference forwards ``UNet(x[b:b+1], time[b:b+1])`` in image order: every image draws
        its own stochastic-depth / expert decisions from Python's ``random`` (unet.py:39, modules.py:35)."""
        t_values = [int(v) for v in time.detach().reshape(-1).tolist()]
        if len(t_values) == 1 and x.shape[0] > 1:
            t_values = t_values * x.shape[0]
        return self._run(x, t_values, plans_per_image=self.draw_plans(x.shape[0]))

    def denoise_step(self, x: torch.Tensor, t: int, coef: "_lib.DdimCoef", noise: Optional[torch.Tensor] = None,
                     out: Optional[torch.Tensor] = None, check_params: bool = True, plans_per_image=None) -> torch.Tensor:
        """One iteration of DDPM.sample (ddpm.py:77-91): eps = UNet(x, t) and the DDIM update fused
        into the network's last kernel.  Writes into ``out`` (default: in place into ``x``)."""
        return self._run(x, [int(t)] * x.shape[0], coef=coef, noise=noise, out=x if out is None else out,
                         check_params=check_params, plans_per_image=plans_per_image)
