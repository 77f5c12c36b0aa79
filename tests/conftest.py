import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import _mol_import  # noqa: E402,F401  (registers the package as `mol_b200`)


@pytest.hookimpl(tryfirst=True)
def pytest_cmdline_main(config):
    """The CPU suite (-m "not gpu") is ~250 host compiles of emulated kernels plus NumPy oracles: spread it over a few
    worker processes when pytest-xdist is installed.  Never for GPU runs (one device, timing-sensitive tests), never when
    -n was given; MOL_TEST_WORKERS=<n> sets the count (0 / 1: serial)."""
    if getattr(config.option, "markexpr", "") != "not gpu" or hasattr(config, "workerinput"):
        return None
    if not config.pluginmanager.hasplugin("xdist") or getattr(config.option, "numprocesses", None) is not None:
        return None
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    env = os.environ.get("MOL_TEST_WORKERS")
    n = int(env) if env not in (None, "") else min(6, cores)
    if n < 2:
        return None
    config.option.numprocesses = n           # (xdist's own hook turns this into dist = "load", tx = n x popen;
    config.option.dist = "load"              #  set here as well in case it has already run)
    config.option.tx = ["popen"] * n
    return None


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
