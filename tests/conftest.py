import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "eo-vae_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree shared library (built by __graft_entry__.build(); never a fallback)."""
    import __graft_entry__ as g
    if not os.path.exists(g.LIB):
        g.build()
    return g.LIB


@pytest.fixture(scope="session")
def cuda(built_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(autouse=True)
def _default_numerics():
    """Every test starts from (and leaves behind) the package's default numeric mode."""
    yield
    import eo_vae
    from eo_vae.settings import default_compute_dtype, default_train_dtype
    eo_vae.set_compute_dtype(default_compute_dtype())
    eo_vae.set_train_dtype(default_train_dtype())
