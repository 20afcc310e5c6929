import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE_DIR = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: imports /root/reference (only present in the build container)")


def have_reference() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "unet.py"))


@pytest.fixture(scope="session")
def reference_modules():
    """The unmodified reference, imported from where it lies (never copied)."""
    if not have_reference():
        pytest.skip("/root/reference not present on this machine")
    sys.path.insert(0, REFERENCE_DIR)
    try:
        import importlib
        mods = {n: importlib.import_module(n) for n in ("unet", "modules", "attention", "sinusoidal", "vae")}
        # ddpm.py builds a default-argument UNet() at import (ddpm.py:16): ~4 s, 385M params.
        mods["ddpm"] = importlib.import_module("ddpm")
    finally:
        sys.path.remove(REFERENCE_DIR)
    return mods
