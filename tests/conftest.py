import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_built():
    from reflib import build_oracle
    return build_oracle()


@pytest.fixture(scope="session")
def sceneprep_lib():
    """The product library; (re)built in-tree when sources are newer (nvcc cross-compiles without a GPU)."""
    from garden_b200.build import build_library
    return build_library()
