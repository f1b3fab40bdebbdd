import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_library():
    """Build (if stale) and return the path of libmsda_b200.so."""
    import __graft_entry__ as entry
    entry.build()
    import vision_instance_seg_b200 as pkg
    return pkg.library_path()
