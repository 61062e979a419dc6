import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the reference (/root/reference or oracle/_ref)")


def pytest_collection_modifyitems(config, items):
    have_ref = any(os.path.isfile(os.path.join(r, "nbm_model", "run_detection.py"))
                   for r in ("/root/reference", os.path.join(ROOT, "oracle", "_ref")))
    skip_ref = pytest.mark.skip(reason="neither /root/reference nor oracle/_ref on this machine")
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
