import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def find_fixture(name):
    """GGUF fixtures are the reference's testdata; they are copied (git-ignored) into
    oracle/_ref/testdata by __graft_entry__.build() so that they travel to the GPU box."""
    for d in (os.path.join(ROOT, "oracle", "_ref", "testdata"), "/root/reference/testdata"):
        p = os.path.join(d, name)
        if os.path.exists(p):
            return p
    return None


@pytest.fixture
def fixture_path():
    def _get(name):
        p = find_fixture(name)
        if p is None:
            pytest.skip(f"fixture {name} not available (run __graft_entry__.build() where /root/reference exists)")
        return p
    return _get
