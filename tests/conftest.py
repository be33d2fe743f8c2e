import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def opts():
    import uvs_b200
    return uvs_b200.default_options()


@pytest.fixture(autouse=True)
def _fused_path_for_every_batch_size(monkeypatch):
    """The library picks the landmark path by batch size (record path below 32 windows, fused linearisation from there on:
    uvs_api.cu).  The parity tests mostly solve single windows, so they pin the fused path unless a test chooses otherwise
    (`landmark_path` parameter of the solve tests, UVS_NO_FUSE / UVS_FUSE_MIN in the environment)."""
    if "UVS_FUSE_MIN" not in os.environ:
        monkeypatch.setenv("UVS_FUSE_MIN", "1")
