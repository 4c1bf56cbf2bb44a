import os
import sys

import pytest

# The GEMM launcher reads this once per process: 2 = also send fp32-partial outputs through the persistent kernel (the
# product default, 1, keeps them on the one-tile kernel), so the GPU suite covers both kernels for every epilogue mode.
os.environ.setdefault("FOLEY_GEMM_PERSIST", "2")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


PKG_DIR = os.path.join(ROOT, "comfyui-hunyuanvideo-foley_b200")


def load_pkg(sub=None):
    """Imports the (hyphen-named) product package under the alias `foley_b200`, like ComfyUI's custom-node
    loader does with importlib; `sub` selects a submodule (e.g. "engine")."""
    import importlib
    import importlib.util
    if "foley_b200" not in sys.modules:
        spec = importlib.util.spec_from_file_location("foley_b200", os.path.join(PKG_DIR, "__init__.py"),
                                                      submodule_search_locations=[PKG_DIR])
        mod = importlib.util.module_from_spec(spec)
        sys.modules["foley_b200"] = mod
        spec.loader.exec_module(mod)
    return importlib.import_module("foley_b200." + sub) if sub else sys.modules["foley_b200"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
