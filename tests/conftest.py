import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_package():
    """Import the product package (its directory name contains a dot, so it is loaded by path
    and registered as `bnr_b200`)."""
    if "bnr_b200" in sys.modules:
        return sys.modules["bnr_b200"]
    pkg_dir = os.path.join(ROOT, "bayesiannetworkregression.jl_b200")
    spec = importlib.util.spec_from_file_location(
        "bnr_b200", os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["bnr_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "golden.npz"))


@pytest.fixture(scope="session")
def bnr():
    so = os.path.join(ROOT, "bayesiannetworkregression.jl_b200", "libbnr.so")
    if not os.path.exists(so):
        # a fresh checkout: build the library first (nvcc cross-compiles sm_100a without a GPU)
        import subprocess
        subprocess.run(["make", "-j4", "-C", ROOT], check=True)
    return load_package()
