import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """Built artefacts are git-ignored: a fresh checkout (or a re-created container) has none.  Build once, up front, instead
    of failing every test that needs the library or silently skipping every test that needs the oracle."""
    lib = os.path.join(ROOT, "rife-ncnn-vulkan_b200", "lib", "librife_b200.so")
    port = os.path.join(ROOT, "oracle", "build", "oracle_rife")
    if os.path.exists(lib) and os.path.exists(port):
        return
    import __graft_entry__ as g
    try:
        g.build()
    except Exception as e:  # the tests that need the artefacts will say what is missing
        sys.stderr.write("conftest: build() failed: %s\n" % e)


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    return g.load_package()
