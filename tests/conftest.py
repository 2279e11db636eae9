import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle32():
    import numpy as np
    from oracle.oracle import Oracle
    return Oracle(np.float32)


@pytest.fixture(scope="session")
def oracle64():
    import numpy as np
    from oracle.oracle import Oracle
    return Oracle(np.float64)


@pytest.fixture(scope="session")
def ref_ext():
    """The reference's own CUDA code compiled for sm_100a (oracle/_ref); skip when not built."""
    import importlib.util
    path = os.path.join(ROOT, "oracle", "_ref", "glic_ref_ext.so")
    if not os.path.isfile(path):
        pytest.skip("oracle/_ref/glic_ref_ext.so not built (needs /root/reference at build time)")
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location("glic_ref_ext", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
