"""CPU: the C-ABI library builds, loads and exports every symbol include/ldmb.h declares; host-side
module API mirrors the reference's (state_dict layout, seeded init, schedule tables).  No compute calls."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from ldm_image_generator_b200 import _lib, build
    build.build_library()
    return _lib.load()


def test_header_symbols_are_exported_and_bound(lib):
    from ldm_image_generator_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "ldmb.h")).read()
    declared = set(re.findall(r"\b(ldmb_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.ldmb_abi_version() == 1


def test_create_fails_loudly_without_a_gpu(lib):
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    assert lib.ldmb_create(0, 0, ctypes.byref(h)) != 0 and not h
    from ldm_image_generator_b200 import UNet, runtime
    m = UNet(3, [1], [32])
    with pytest.raises(runtime.LdmbError):
        m(torch.zeros(1, 3, 8, 8), torch.zeros(1, dtype=torch.long))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ldm_image_generator_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, f"{f} mentions the oracle"


def test_module_api_layout_matches_oracle_key_templates():
    """Constructor signatures + state_dict layout of the drop-in modules."""
    from ldm_image_generator_b200 import DDPM, Decoder, Encoder, UNet
    from oracle import restate as R
    cfg = R.UNetCfg(input_channels=8, stages=(1, 2, 2), channels=(32, 64, 128))
    sd = R.make_unet_state(cfg, 7)
    m = UNet(cfg.input_channels, list(cfg.stages), list(cfg.channels), cfg.stem_size)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    m.load_state_dict(sd, strict=True)
    d = DDPM(model=m)
    assert all(k.startswith("model.") for k in d.state_dict()) and len(d.state_dict()) == len(sd)
    assert {k: tuple(v.shape) for k, v in Decoder().state_dict().items()} == \
        {k: tuple(v.shape) for k, v in R.make_decoder_state(R.DecoderCfg()).items()}
    assert {k: tuple(v.shape) for k, v in Encoder().state_dict().items()} == \
        {k: tuple(v.shape) for k, v in R.make_encoder_state(R.EncoderCfg()).items()}


def test_host_schedule_is_bit_exact():
    from ldm_image_generator_b200 import DDPM, UNet
    from oracle import restate as R
    d = DDPM(model=UNet(3, [1], [32]))
    fix = torch.load(os.path.join(ROOT, "tests", "golden", "schedule.pt"))
    assert torch.equal(d.beta, fix["beta"])
    assert torch.equal(torch.cumprod(1 - d.beta, dim=0), fix["alpha"])
    for n in (20, 50, 1000):
        assert [(int(a), int(b)) for a, b in d.timesteps(n)] == R.step_pairs(R.linear_steps(1000, n))
    assert d.timesteps(3, [5, 9, 100]) == [(100, 9), (9, 5), (5, 0)]
    with pytest.raises(TypeError):
        d.timesteps(3, "cosine")
    co, sigma = d.ddim_scalars(fix["alpha"], 999, 978, 0)
    want = R.ddim_coefficients(fix["alpha"], 999, 978, 0.0)
    for k in ("c_eps_in", "c_div", "c_x0", "c_eps_out", "sigma"):
        assert getattr(co, k) == float(want[k])


def test_host_plan_and_tables_match_oracle():
    import random
    from ldm_image_generator_b200 import UNet
    from ldm_image_generator_b200.sinusoidal import position_table, time_table
    from oracle import restate as R
    m = UNet()
    for training in (False, True):
        m.train(training)
        random.seed(0); a = m.draw_plan(); sa = random.getstate()
        random.seed(0); b = R.draw_plan(36, training)
        assert [tuple(p) for p in a] == b and random.getstate() == sa
    assert torch.equal(position_table(128, 32, 32), R.position_table(128, 32, 32))
    t = torch.tensor([0, 20, 999])
    assert torch.equal(time_table(256, t), R.time_table(256, t))


def test_plan_replay_in_c_consumes_python_random_exactly():
    """UNet.draw_plans(n) (ldmb_host_draw_plans replaying the raw MT19937 stream: the per-image sampler's images x steps
    decisions) == n calls of draw_plan() == the oracle's, with `random` left in the same state; mixed train/eval blocks,
    a generator position in the middle of its 624-word block, and a stream long enough to cross regenerations."""
    import random
    import numpy as np
    from ldm_image_generator_b200 import UNet
    from oracle import restate as R
    m = UNet(8, [2, 3], [32, 64])
    nblk = len(m.blocks_in_execution_order())
    for mode in ("eval", "train", "mixed"):
        m.train(mode != "eval")
        if mode == "mixed":
            for b in m.blocks_in_execution_order()[::2]:
                b.eval()
        for seed, n in ((0, 1), (7, 300)):
            random.seed(seed); [random.random() for _ in range(seed)]
            fast = m.draw_plans(n); end = random.getstate()
            random.seed(seed); [random.random() for _ in range(seed)]
            slow = np.array([m.draw_plan() for _ in range(n)], dtype=np.int32)
            assert UNet._replay_checked is True
            assert fast.shape == (n, nblk, 3) and np.array_equal(fast, slow) and random.getstate() == end
            if mode != "mixed":
                random.seed(seed); [random.random() for _ in range(seed)]
                assert [tuple(p) for p in fast[-1].tolist()] == [R.draw_plan(nblk, mode == "train") for _ in range(n)][-1]


def test_modules_deepcopy_and_pickle_without_a_handle():
    """The reference modules can be deep-copied and torch.save'd; the ctypes handle is runtime state and is dropped."""
    import copy
    import io
    from ldm_image_generator_b200 import Decoder, UNet
    u = UNet(8, [1, 1], [64, 128])
    u._handle, u._film = object(), (1, 2, 3)           # stand-ins for live runtime state
    v = copy.deepcopy(u)
    assert v._handle is None and v._film is None and u._handle is not None
    buf = io.BytesIO()
    u._handle = None
    torch.save(u, buf)
    d = copy.deepcopy(Decoder(channels=[64, 32], stages=[1, 1]))
    assert d._handle is None
    u.invalidate_weights(); d.invalidate_weights()      # no handle yet: no-ops


def test_reference_staging_recipe_matches_the_reference():
    """oracle/make_ref.py stages the unmodified reference byte for byte (build container only)."""
    import hashlib
    import json
    import os
    import sys
    if not os.path.isfile("/root/reference/unet.py"):
        pytest.skip("/root/reference not present on this machine")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "oracle"))
    try:
        import make_ref
    finally:
        sys.path.pop(0)
    d = make_ref.stage()
    man = json.load(open(os.path.join(d, "MANIFEST.json")))["files"]
    assert set(make_ref.FILES) == set(man)
    for f, digest in man.items():
        assert hashlib.sha256(open(os.path.join("/root/reference", f), "rb").read()).hexdigest() == digest
    assert make_ref.available() == d
