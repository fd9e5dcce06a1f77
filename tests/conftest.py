import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_pkg():
    """import the package whose directory name contains a dot"""
    if "welle_io_b200" in sys.modules:
        return sys.modules["welle_io_b200"]
    d = os.path.join(ROOT, "welle.io_b200")
    spec = importlib.util.spec_from_file_location("welle_io_b200", os.path.join(d, "__init__.py"), submodule_search_locations=[d])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["welle_io_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


@pytest.fixture(scope="session")
def oracle():
    from oracle.bind import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle.bind import Ref
    if not Ref.available():
        if os.path.isdir("/root/reference/src"):
            from oracle.bind import build_ref
            build_ref()
        else:
            pytest.skip("oracle/_ref/libwelle_ref.so not built (needs /root/reference)")
    return Ref()
