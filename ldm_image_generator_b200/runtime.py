"""Thin object wrapper over the C ABI: one ``Handle`` per module instance per device.

PyTorch is used here only as plumbing -- it owns the device tensors and the current
stream; every arithmetic operator of the path runs inside libldmb200.so.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Iterable, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

PRECISIONS = {"bf16": _lib.BF16, "fp32": _lib.FP32_VALIDATE}


class LdmbError(RuntimeError):
    pass


def default_precision() -> str:
    p = os.environ.get("LDMB_PRECISION", "bf16").lower()
    if p not in PRECISIONS:
        raise ValueError(f"LDMB_PRECISION must be one of {sorted(PRECISIONS)}, got {p!r}")
    return p


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise LdmbError(
            f"{what} is on {t.device}: this implementation runs only on a CUDA device (sm_100a); "
            "there is no CPU fallback. Move the module and its inputs to 'cuda'.")


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def f32c(t: torch.Tensor) -> torch.Tensor:
    """fp32 contiguous view/copy (host code only; no arithmetic)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


class Handle:
    def __init__(self, device: torch.device, precision: str):
        if device.type != "cuda":
            raise LdmbError(f"cannot create a libldmb200 handle on {device}: CUDA only, no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device())
        self.precision = precision
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.cuda.current_stream()      # make sure the primary context exists before the library touches it
        h = C.c_void_p()
        rc = self.lib.ldmb_create(self.device.index, PRECISIONS[precision], C.byref(h))
        if rc != 0 or not h:
            raise LdmbError(f"ldmb_create(device={self.device.index}) failed with status {_lib.STATUS.get(rc, rc)} "
                            "(needs an sm_100a GPU)")
        self.h = h
        self._param_keys: Dict[str, Tuple] = {}
        self._keepalive = []

    # ------------------------------------------------------------------ plumbing
    def check(self, rc: int) -> None:
        if rc != 0:
            msg = self.lib.ldmb_last_error(self.h)
            raise LdmbError(f"libldmb200: {_lib.STATUS.get(rc, rc)}: {msg.decode() if msg else ''}")

    def close(self) -> None:
        if getattr(self, "h", None):
            self.lib.ldmb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(self.lib.ldmb_launch_count(self.h))

    def set_force_simt(self, on: bool) -> None:
        self.check(self.lib.ldmb_set_force_simt(self.h, int(on)))

    def set_deterministic(self, on: bool) -> None:
        """Bit-reproducible bf16 results (no split-K slices, no concurrent conv branch); see ldmb.h."""
        self.check(self.lib.ldmb_set_deterministic(self.h, int(on)))

    def set_use_graphs(self, on: bool) -> None:
        self.check(self.lib.ldmb_set_use_graphs(self.h, int(on)))

    def device_fault(self) -> int:
        """The watchdog word, read on the device: synchronises the current stream."""
        return int(self.lib.ldmb_check_device_fault(self.h, stream_ptr(self.device)))

    def raise_on_fault(self, synchronize: bool = False) -> None:
        """Raise ``LdmbError`` if a tcgen05 pipeline watchdog has fired on this handle -- results produced since are
        garbage.  Without ``synchronize`` the host-mapped mirror is polled (no stall; sees kernels that have run)."""
        code = self.device_fault() if synchronize else int(self.lib.ldmb_poll_device_fault(self.h))
        if code != 0:
            raise LdmbError(f"libldmb200: kernel fault: a tcgen05 pipeline watchdog fired (code {code}); "
                            "outputs computed on this handle since then are invalid")

    PROFILE_CLASSES = ("gemm_ffn_ab", "gemm_ffn_c", "gemm_qkv", "gemm_encodings", "gemm_level_change", "grouped_conv3x3",
                       "vae_conv3x3", "vae_gemm", "gemm_cuda_core", "channelnorm_film", "window_attention", "stem_final", "other")
    TENSOR_CLASSES = PROFILE_CLASSES[:9]

    def skip_classes(self, names) -> None:
        """Debug (tools/ablate.py): drop the launches of these classes (results become garbage)."""
        mask = 0
        for n in names:
            mask |= 1 << self.PROFILE_CLASSES.index(n)
        self.check(self.lib.ldmb_debug_skip_classes(self.h, mask))


    def profile_begin(self) -> None:
        self.check(self.lib.ldmb_profile_begin(self.h))

    def profile_end(self):
        """{class: (ms, work, launches)}; work = FLOPs for GEMM/conv classes, bytes for HBM-bound ones."""
        n = len(self.PROFILE_CLASSES)
        ms, work, cnt = (C.c_double * n)(), (C.c_double * n)(), (C.c_int64 * n)()
        self.check(self.lib.ldmb_profile_end(self.h, ms, work, cnt))
        return {k: (ms[i], work[i], int(cnt[i])) for i, k in enumerate(self.PROFILE_CLASSES)}

    def sync_params(self, loader, items: Iterable[Tuple[str, torch.Tensor]]) -> int:
        """Upload every state_dict entry whose storage or version changed since the last call."""
        n = 0
        st = stream_ptr(self.device)
        for name, t in items:
            key = (t.data_ptr(), t._version, tuple(t.shape), t.dtype)
            if self._param_keys.get(name) == key:
                continue
            _require_cuda(t, f"parameter {name}")
            src = f32c(t.detach())
            shape = (C.c_int64 * max(src.dim(), 1))(*src.shape)
            self.check(loader(name.encode(), src.data_ptr(), shape, src.dim(), st))
            self._param_keys[name] = key
            n += 1
        return n

    # ------------------------------------------------------------------ UNet
    def unet_configure(self, input_channels: int, stages: Sequence[int], channels: Sequence[int], stem_size: int) -> None:
        cfg = _lib.UNetConfig()
        cfg.input_channels, cfg.num_levels, cfg.stem_size = input_channels, len(stages), stem_size
        if len(stages) > _lib.LDMB_MAX_LEVELS or len(stages) != len(channels):
            raise LdmbError("stages/channels must have equal length <= 8")
        for i, (b, c) in enumerate(zip(stages, channels)):
            cfg.blocks[i], cfg.channels[i] = int(b), int(c)
        self.check(self.lib.ldmb_unet_configure(self.h, C.byref(cfg)))

    def unet_load(self, items) -> int:
        return self.sync_params(lambda *a: self.lib.ldmb_unet_load_param(self.h, *a), items)

    def unet_set_position_table(self, level: int, pe_chw: torch.Tensor) -> None:
        pe = f32c(pe_chw.cpu())
        Cc, Hl, Wl = pe.shape
        self.check(self.lib.ldmb_unet_set_position_table(self.h, level, pe.data_ptr(), Cc, Hl, Wl, stream_ptr(self.device)))

    def unet_precompute_film(self, H: int, W: int, te_tables: Sequence[torch.Tensor]) -> None:
        te_ptrs = (C.c_void_p * len(te_tables))(*[t.data_ptr() for t in te_tables])
        self.check(self.lib.ldmb_unet_precompute_film(self.h, H, W, te_tables[0].shape[0], te_ptrs, stream_ptr(self.device)))

    def unet_forward(self, x: torch.Tensor, out: torch.Tensor, t_index, te_tables: Sequence[torch.Tensor], plan,
                     coef: Optional[_lib.DdimCoef] = None, noise: Optional[torch.Tensor] = None, per_image: bool = False) -> None:
        """plan: [(skip, e1, e2)] per block, or with ``per_image`` [[(skip, e1, e2)] per image] per block."""
        B, _, H, W = x.shape
        if isinstance(te_tables, int):       # FiLM tables precomputed for this many timesteps (unet_precompute_film)
            n_t, te_ptrs = te_tables, None
        else:
            n_t = te_tables[0].shape[0]
            te_ptrs = (C.c_void_p * len(te_tables))(*[t.data_ptr() for t in te_tables])
        ti = (C.c_int32 * B)(*t_index)
        if per_image:
            keep = np.ascontiguousarray(np.asarray(plan, dtype=np.int32))
            if keep.ndim != 3 or keep.shape[1] != B or keep.shape[2] != 3:
                raise LdmbError("per-image plan must be [n_blocks][B][3]")
            pl = keep.ctypes.data_as(C.POINTER(C.c_int32))
        else:
            flat = [int(v) for row in plan for v in row]
            pl = (C.c_int32 * len(flat))(*flat)
        fn = self.lib.ldmb_unet_forward_per_image if per_image else self.lib.ldmb_unet_forward
        self.check(fn(
            self.h, x.data_ptr(), out.data_ptr(), B, H, W, ti, n_t, te_ptrs, pl,
            C.byref(coef) if coef is not None else None,
            noise.data_ptr() if noise is not None else None, stream_ptr(self.device)))

    # ------------------------------------------------------------------ VAE
    def vae_configure(self, which: int, image_channels: int, latent_channels: int, channels: Sequence[int],
                      stages: Sequence[int]) -> None:
        cfg = _lib.VaeConfig()
        cfg.image_channels, cfg.latent_channels, cfg.num_levels = image_channels, latent_channels, len(channels)
        if len(channels) > _lib.LDMB_MAX_LEVELS or len(stages) != len(channels):
            raise LdmbError("stages/channels must have equal length <= 8")
        for i, (c, b) in enumerate(zip(channels, stages)):
            cfg.channels[i], cfg.blocks[i] = int(c), int(b)
        self.check(self.lib.ldmb_vae_configure(self.h, which, C.byref(cfg)))

    def vae_load(self, which: int, items) -> int:
        return self.sync_params(lambda *a: self.lib.ldmb_vae_load_param(self.h, which, *a), items)

    def vae_decode(self, z: torch.Tensor, img: Optional[torch.Tensor], img_u8: Optional[torch.Tensor]) -> None:
        B, _, hl, wl = z.shape
        self.check(self.lib.ldmb_vae_decode(self.h, z.data_ptr(), img.data_ptr() if img is not None else None,
                                            img_u8.data_ptr() if img_u8 is not None else None, B, hl, wl,
                                            stream_ptr(self.device)))

    def vae_encode(self, img: torch.Tensor, z: torch.Tensor) -> None:
        B, _, H, W = img.shape
        self.check(self.lib.ldmb_vae_encode(self.h, img.data_ptr(), z.data_ptr(), B, H, W, stream_ptr(self.device)))

    # ------------------------------------------------------------------ kernel-level (tests / bench)
    def gemm(self, A, Wt, bias, out, M, N, K, out_f32=0, act=0, force_simt=False) -> None:
        self.check(self.lib.ldmb_gemm(self.h, A.data_ptr(), Wt.data_ptr(), bias.data_ptr() if bias is not None else None,
                                      out.data_ptr(), M, N, K, out_f32, act, int(force_simt), stream_ptr(self.device)))

    def conv3x3(self, x, Wt, bias, out, B, H, W, Cc, N, act=0, force_simt=False) -> None:
        self.check(self.lib.ldmb_conv3x3(self.h, x.data_ptr(), Wt.data_ptr(), bias.data_ptr() if bias is not None else None,
                                         out.data_ptr(), B, H, W, Cc, N, act, int(force_simt), stream_ptr(self.device)))

    def mlp_fused(self, xm, w_ab, b_ab, w_c, b_c, x, M, Cc, e1, e2) -> None:
        self.check(self.lib.ldmb_mlp_fused(self.h, xm.data_ptr(), w_ab.data_ptr(), b_ab.data_ptr(), w_c.data_ptr(), b_c.data_ptr(),
                                           x.data_ptr(), M, Cc, e1, e2, stream_ptr(self.device)))

    def mlp_fused_attn(self, xm, w_ab, b_ab, w_c, b_c, att, x, M, Cc, e1, e2) -> None:
        self.check(self.lib.ldmb_mlp_fused_attn(self.h, xm.data_ptr(), w_ab.data_ptr(), b_ab.data_ptr(), w_c.data_ptr(), b_c.data_ptr(),
                                                att.data_ptr(), att.stride(-2), x.data_ptr(), M, Cc, e1, e2, stream_ptr(self.device)))

    def normconv(self, x, film, xm, w_packed, bias, B, H, W, Cc) -> None:
        self.check(self.lib.ldmb_normconv(self.h, x.data_ptr(), film.data_ptr(), xm.data_ptr(), w_packed.data_ptr(), bias.data_ptr(),
                                          B, H, W, Cc, stream_ptr(self.device)))

    def grouped_conv3x3(self, xm, w_packed, bias, x, B, H, W, Cc, force_generic=False) -> None:
        self.check(self.lib.ldmb_grouped_conv3x3(self.h, xm.data_ptr(), w_packed.data_ptr(), bias.data_ptr(), x.data_ptr(),
                                                 B, H, W, Cc, int(force_generic), stream_ptr(self.device)))

    def window_attention(self, qkv, xm, b_in, att, B, H, W, Cc, win_h, win_w, shift, force_simt=False) -> None:
        self.check(self.lib.ldmb_window_attention(self.h, qkv.data_ptr(), xm.data_ptr(), b_in.data_ptr(), att.data_ptr(),
                                                  att.stride(-2), B, H, W, Cc, win_h, win_w, shift, int(force_simt),   # 0 tcgen05, 1 CUDA cores, 2 mma.sync
                                                  stream_ptr(self.device)))

    def channelnorm_film(self, x, film, out, M, Cc, HW) -> None:
        self.check(self.lib.ldmb_channelnorm_film(self.h, x.data_ptr(), film.data_ptr(), out.data_ptr(), M, Cc, HW,
                                                  stream_ptr(self.device)))
