"""pytest configuration: registers the `gpu` marker (tests that need a B200 and call through the C-ABI)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


def _have_gpu() -> bool:
    """True when the in-tree library loads and sees a CUDA device (the only way the -m gpu tests can run)"""
    try:
        import compute_war_b200 as cw
        return cw.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if any("gpu" in it.keywords for it in items) and not _have_gpu():
        skip = pytest.mark.skip(reason="no CUDA device (or libcomputewar.so not built): GPU parity tests need the B200 box")
        for it in items:
            if "gpu" in it.keywords:
                it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import json
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    import numpy as np
    with open(os.path.join(here, "skein_kat.json")) as f:
        kat = json.load(f)
    with open(os.path.join(here, "ref_outputs.json")) as f:
        ref = json.load(f)
    blob = np.fromfile(os.path.join(here, "corpus_sample.bin"), np.uint8)
    with open(os.path.join(here, "fast_mode_outputs.json")) as f:
        fast = json.load(f)  # the engine's own fast-mode parses, pinned (make_fast_mode_golden.py)
    return {"kat": kat, "ref": ref, "blob": blob, "fast": fast}
