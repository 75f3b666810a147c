import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _has_gpu() -> bool:
    try:
        from ragnar_b200 import cabi

        return cabi.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def cabi():
    """The ctypes C-ABI binding, initialised on cuda:0 (GPU tests only)."""
    from ragnar_b200 import cabi as _cabi

    _cabi.init(int(os.environ.get("LOCAL_RANK", "0")))
    return _cabi


@pytest.fixture(scope="session")
def rg(cabi):
    """The compiled `ragnar` module, initialised."""
    import ragnar_b200

    mod = ragnar_b200.load()
    mod.Initialize()
    return mod


@pytest.fixture(scope="session")
def port():
    import oracle

    return oracle.port
