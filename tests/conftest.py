import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG_NAME = "radad-retrievalaugmenteddeepfakeaudiodetection_b200"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (hyphenated directory name -> importlib)."""
    return importlib.import_module(PKG_NAME)


@pytest.fixture(scope="session")
def oracle():
    return importlib.import_module("oracle.flat_oracle")


class Cfg:
    """Minimal stand-in for the reference's Config attribute bag (config.py:18-115): only the keys the
    hot path reads."""

    def __init__(self, path, index_type="L2", top_k=5, **kw):
        self.vector_db_path = str(path)
        self.vector_db_index_type = index_type
        self.top_k = top_k
        self.use_float16 = False
        self.vector_add_batch_size = 10000
        self.vector_db_nprobe = 32
        for k, v in kw.items():
            setattr(self, k, v)


@pytest.fixture
def make_cfg(tmp_path):
    def _mk(index_type="L2", top_k=5, **kw):
        return Cfg(tmp_path / f"vdb_{index_type}_{len(os.listdir(tmp_path))}", index_type, top_k, **kw)
    return _mk
