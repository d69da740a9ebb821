import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gcn-bmp_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the C-ABI library is a build product (git-ignored): compile it when it is missing (a fresh checkout); a stale or
    # unbuildable library is left to fail loudly in the tests that load it
    import importlib.util
    spec = importlib.util.spec_from_file_location("_bmp_build", os.path.join(ROOT, "gcn-bmp_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if not os.path.exists(mod.OUT):
        try:
            mod.build()
        except Exception as exc:      # pragma: no cover
            sys.stderr.write("conftest: could not build libgcnbmp.so: %s\n" % exc)


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
