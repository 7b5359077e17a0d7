import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # Safety net: (re)build libsnsde.so when it is missing or older than its sources (a no-op otherwise; the build
    # is the in-tree nvcc recipe of __graft_entry__.build()).  A failure here is reported by the tests themselves.
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("snsde_build", ROOT / "stable-neural-sdes_b200" / "build.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    except Exception as exc:                     # noqa: BLE001
        print(f"[conftest] could not build libsnsde.so: {exc}", file=sys.stderr)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
