import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_cases():
    from oracle.make_golden import load_case

    paths = sorted((ROOT / "tests" / "golden").glob("*.npz"))
    assert paths, "no golden fixtures found"
    return {p.stem: load_case(p) for p in paths}
