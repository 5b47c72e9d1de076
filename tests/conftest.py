import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    # a gpu-marked test on a box without a GPU is a configuration error, not a skip: fail loudly
    # only when the user explicitly selected them (-m gpu); the default CPU run deselects via -m "not gpu".
    pass
