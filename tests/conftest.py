import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib():
    import bzip2_b200
    if not os.path.exists(bzip2_b200.LIB_PATH):
        bzip2_b200.build_library()
    return bzip2_b200.load()


_engines = {}


@pytest.fixture(scope="session")
def engine_for():
    """Engines are expensive (several GB of HBM); share one per level across the session."""
    import bzip2_b200

    def get(level):
        if level not in _engines:
            # keep at most two alive
            while len(_engines) >= 2:
                _engines.pop(next(iter(_engines))).close()
            _engines[level] = bzip2_b200.Engine(level=level)
        return _engines[level]

    yield get
    for e in _engines.values():
        e.close()
    _engines.clear()
