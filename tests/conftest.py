import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import __graft_entry__ as entry  # noqa: E402

entry.load_package()             # registers `bonnie32_b200`


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure)."""
    entry.build_oracle()
    from oracle import oracle as orc
    return orc


@pytest.fixture(scope="session")
def ctx():
    """One CUDA context for the whole GPU session; fails loudly without the extension or a GPU."""
    import bonnie32_b200 as pkg
    c = pkg.Context(0)
    yield c
    c.close()
