import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        import spinoza_b200 as sb
        return sb.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # A GPU test on a box without a GPU is skipped, never silently passed on a CPU fallback.
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
