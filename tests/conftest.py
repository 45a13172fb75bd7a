import importlib.util
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def load_package():
    """The product directory is `minbpe-cc_b200/` (hyphen): import it as module `minbpe_cc_b200`."""
    if "minbpe_cc_b200" in sys.modules:
        return sys.modules["minbpe_cc_b200"]
    spec = importlib.util.spec_from_file_location("minbpe_cc_b200", os.path.join(ROOT, "minbpe-cc_b200", "__init__.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["minbpe_cc_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def pkg():
    mod = load_package()
    if not os.path.exists(mod.LIB_PATH):
        subprocess.check_call([sys.executable, os.path.join(ROOT, "minbpe-cc_b200", "build.py")])
    mod.lib()
    return mod


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        return json.load(f)


def golden_data(name):
    with open(os.path.join(GOLDEN, "data", name), "rb") as f:
        return f.read()


def train_cases(manifest, max_input_bytes=None, slow_first=False):
    out = []
    for name, e in sorted(manifest["train"].items()):
        if e["rc"] != 0:
            continue
        if max_input_bytes is not None and os.path.getsize(os.path.join(GOLDEN, "data", e["input"])) > max_input_bytes:
            continue
        out.append((name, e))
    return out
