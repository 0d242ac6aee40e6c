import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)


def load_package():
    """Import the product package (directory `sph-bvf_b200`, hyphenated like the reference repo)."""
    name = "sphbvf_b200"
    if name in sys.modules:
        return sys.modules[name]
    path = os.path.join(ROOT, "sph-bvf_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
