#!/usr/bin/env python
"""Recipe for ``oracle/_ref``: the UNMODIFIED reference, staged where the GPU box can import it.

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference (uthree/ldm-image-generator) is a flat directory of Python
scripts with no build system: there is nothing to compile, so the "build" of ``oracle/_ref`` is a byte-for-byte
staging copy of the modules of the sampling path from where they lie under ``/root/reference`` into the
git-ignored ``oracle/_ref/`` (it travels to the GPU box like the built ``.so``; it never enters history, and no
reference source is committed).  ``MANIFEST.json`` records the sha256 of every staged file so a consumer can tell
that what it imports is what the reference ships.

Consumers (the only ones allowed): ``bench.py --impl reference`` / the ``cpu_baseline`` and ``gpu_baseline`` legs
(the reference timed on the box's host cores and, as PyTorch eager, on the B200), and ``tests/``.  The product
package never imports it.

    python oracle/make_ref.py            # no-op (keeps what is staged) when /root/reference is absent
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
# the modules sample_ldm.py / sample_ddpm.py import (sample_ldm.py:1-2, ddpm.py:6, unet.py:5-7, vae.py) + the scripts
FILES = ["ddpm.py", "unet.py", "modules.py", "attention.py", "sinusoidal.py", "vae.py", "sample_ldm.py", "sample_ddpm.py"]


def sha256(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def stage() -> str | None:
    """Stage the reference; returns the directory, or None when there is neither a source nor a staged copy."""
    if not os.path.isfile(os.path.join(SRC, "unet.py")):
        return DST if os.path.isfile(os.path.join(DST, "MANIFEST.json")) else None
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
        manifest[f] = sha256(os.path.join(DST, f))
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "files": manifest}, fh, indent=1)
    return DST


def available() -> str | None:
    """The staged directory if every file matches its recorded digest."""
    mp = os.path.join(DST, "MANIFEST.json")
    if not os.path.isfile(mp):
        return None
    files = json.load(open(mp))["files"]
    for f, digest in files.items():
        p = os.path.join(DST, f)
        if not os.path.isfile(p) or sha256(p) != digest:
            return None
    return DST


def import_reference():
    """{'ddpm': module, 'unet': ..., 'vae': ...} of the staged reference, or None.  ``import ddpm`` builds the
    default-argument UNet() (ddpm.py:16, ~4 s, 385.7 M parameters) exactly as the reference does."""
    d = available()
    if d is None:
        return None
    import importlib
    sys.path.insert(0, d)
    try:
        for n in ("unet", "modules", "attention", "sinusoidal", "vae", "ddpm"):
            if n in sys.modules and not getattr(sys.modules[n], "__file__", "").startswith(d):
                del sys.modules[n]
        return {n: importlib.import_module(n) for n in ("unet", "modules", "attention", "sinusoidal", "vae", "ddpm")}
    finally:
        sys.path.remove(d)


if __name__ == "__main__":
    print(stage())
