#!/usr/bin/env python
"""Checksums of the reference's image preprocessing (dataset.py:119-175 with an Identity encoder) on the synthetic
sources of tests/test_latent_cache.py -> tests/golden/latent_cache.json.  Run in the build container (needs /root/reference)."""
import json
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.test_latent_cache import SIZE, make_images, sha  # noqa: E402

sys.path.insert(0, "/root/reference")
import dataset as ref_dataset  # noqa: E402

with tempfile.TemporaryDirectory() as tmp:
    src = make_images(os.path.join(tmp, "src"))
    cache = os.path.join(tmp, "cache") + "/"
    ds = ref_dataset.LatentImageDataset([src], cache_dir=cache, size=SIZE, max_len=None, encoder=torch.nn.Identity(), n_workers=1)
    out = {"size": SIZE, "sha": {os.path.relpath(p, src): sha(torch.load(os.path.join(cache, f"{i}.pt")))
                                 for i, p in enumerate(ds.image_path_list)}}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "latent_cache.json"), "w"), indent=1)
print(out)
