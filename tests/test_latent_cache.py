"""Latent pre-encoding (dataset.py:99-189 of the reference): host preprocessing bit-exact against the reference's own
LatentImageDataset (build container) and against committed checksums (everywhere); GPU encode against the oracle."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest
import torch
from PIL import Image

from ldm_image_generator_b200 import latent_cache as LC

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "latent_cache.json")
REF = "/root/reference"
SIZE = 64


def make_images(root):
    """Deterministic synthetic sources: landscape / portrait / small / large, PNG (top level) and JPG (nested)."""
    rng = np.random.RandomState(7)
    os.makedirs(os.path.join(root, "sub"), exist_ok=True)
    specs = [("a.png", 96, 40), ("b.png", 20, 30), ("sub/c.jpg", 50, 128), ("sub/d.jpg", 64, 64), ("e.png", 200, 17)]
    for name, w, h in specs:
        arr = rng.randint(0, 256, size=(h, w, 3), dtype=np.uint8)
        arr[: h // 2] = (arr[: h // 2] // 4) * 4            # some structure so the blur matters
        Image.fromarray(arr, "RGB").save(os.path.join(root, name), quality=95)
    return root


def sha(t):
    return hashlib.sha256(np.ascontiguousarray(t.numpy()).tobytes()).hexdigest()[:16]


def test_preprocessing_matches_committed_checksums(tmp_path):
    """Checksums made by the unmodified reference (oracle/gen_latent_cache_golden.py); keyed by file because glob order,
    and with it the image the reference's default max_len=-1 drops, depends on the file system."""
    src = make_images(str(tmp_path / "src"))
    want = json.load(open(GOLDEN))["sha"]
    paths = LC.list_images([src], max_len=None)
    assert sorted(os.path.relpath(p, src) for p in paths) == sorted(want)
    assert LC.list_images([src]) == paths[:-1]              # dataset.py:107 with the default max_len=-1
    n = LC.encode_image_folder([src], str(tmp_path / "cache"), torch.nn.Identity(), size=SIZE, batch=2, device="cpu", paths=paths)
    assert n == len(want)
    for i, p in enumerate(paths):
        z = torch.load(str(tmp_path / "cache" / f"{i}.pt"))
        assert tuple(z.shape) == (1, 3, SIZE, SIZE) and z.dtype == torch.float32
        assert sha(z) == want[os.path.relpath(p, src)], p
    ds = LC.LatentCache(str(tmp_path / "cache"))
    assert len(ds) == n and tuple(ds[1].shape) == (3, SIZE, SIZE)


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference is only present in the build container")
def test_preprocessing_matches_the_reference_dataset(tmp_path):
    src = make_images(str(tmp_path / "src"))
    sys.path.insert(0, REF)
    try:
        import dataset as ref_dataset                     # the unmodified reference module
    finally:
        sys.path.remove(REF)
    ref_cache, our_cache = str(tmp_path / "ref_cache") + "/", str(tmp_path / "our_cache")
    ref_dataset.LatentImageDataset([src], cache_dir=ref_cache, size=SIZE, encoder=torch.nn.Identity(), n_workers=1)
    n = LC.encode_image_folder([src], our_cache, torch.nn.Identity(), size=SIZE, batch=3, device="cpu")
    assert n == len(os.listdir(ref_cache)) > 0
    for i in range(n):
        assert torch.equal(torch.load(os.path.join(ref_cache, f"{i}.pt")), torch.load(os.path.join(our_cache, f"{i}.pt"))), i


@pytest.mark.gpu
def test_batched_gpu_encode_writes_the_reference_cache_format(tmp_path):
    from oracle import restate as R
    from tests.gpu_util import assert_no_fault, build_encoder
    src = make_images(str(tmp_path / "src"))
    cfg = R.EncoderCfg(channels=(8, 16, 32, 64))
    sd = R.make_encoder_state(cfg, 5)
    enc = build_encoder(cfg, sd, "fp32")
    n = LC.encode_image_folder([src], str(tmp_path / "cache"), enc, size=SIZE, batch=3, device="cuda")
    paths = LC.list_images([src])
    assert n == len(paths) == 4
    for i, p in enumerate(paths):
        z = torch.load(str(tmp_path / "cache" / f"{i}.pt"))
        assert z.is_cuda and tuple(z.shape) == (1, 8, SIZE // 8, SIZE // 8)
        with Image.open(p) as im:
            x = torch.from_numpy(LC.preprocess_image(im, SIZE))[None]
        assert R.rel_l2(z.cpu(), R.encoder_forward(sd, cfg, x)) < 1e-5
    assert_no_fault(enc)
