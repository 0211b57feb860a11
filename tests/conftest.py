"""pytest configuration: `gpu` marker, import paths, golden-fixture helpers."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "sciml-pde_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with `-m gpu` on the GPU box")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_spectral():
    return np.load(GOLDEN / "spectral_small.npz")


@pytest.fixture(scope="session")
def golden_models():
    return np.load(GOLDEN / "models_small.npz")


@pytest.fixture(scope="session")
def golden_cfg1():
    import json

    return json.loads((GOLDEN / "cfg1_meta.json").read_text()), np.load(GOLDEN / "cfg1_samples.npz")
