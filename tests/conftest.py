"""pytest config: registers the ``gpu`` marker; GPU tests are skipped when no CUDA device exists."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# no Whisper checkpoint exists offline: the tests opt in to the name-seeded random weights the goldens were made with
# (aga_b200.whisper_model.load_model raises without this, see test_host_mirror.py::test_load_model_needs_checkpoint_or_opt_in)
os.environ.setdefault("AGA_ALLOW_RANDOM_INIT", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
