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


# Tuning switches of the engine change how a circuit is scheduled or which kernel variant runs, never
# what it computes; the suite pins them to their defaults so that shape-specific assertions do not depend on the caller's shell.
TUNING_ENV = ("SPZ_TILE_V3", "SPZ_TILE_LMIN", "SPZ_TILE_SELECT", "SPZ_DIST_WINDOW", "SPZ_TILE_MIN_OPS", "SPZ_DIST_FUSE_GATE", "SPZ_XG_CTAS")


@pytest.fixture(autouse=True)
def _default_tuning_env(monkeypatch):
    if os.environ.get("SPZ_TEST_KEEP_ENV") == "1":  # e.g. the whole GPU suite under SPZ_TILE_V3=0
        return
    for k in TUNING_ENV:
        monkeypatch.delenv(k, raising=False)
