import os
import sys

import pytest

# poison the device output arena before every decode so that a byte a kernel failed to write (or read too
# early) can never be masked by stale data of an earlier call
os.environ.setdefault("SDZ_POISON", "1")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sd-zlib_b200", "host"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

FIXTURES = os.path.join(ROOT, "tests", "golden", "ref_fixtures")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def fixture_bytes(name):
    with open(os.path.join(FIXTURES, name), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def fx():
    return fixture_bytes
