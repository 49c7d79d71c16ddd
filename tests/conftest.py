import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def pytest_sessionfinish(session, exitstatus):
    """Per-stage parity numbers of the GPU tests (max / mean |logit diff|, agreement, near-tie counts and margins) -> gpurun_out/parity_r02.json."""
    mod = sys.modules.get("tests.parity_utils")
    if mod is None or not getattr(mod, "RECORDS", None):
        return
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_r02.json"), "w") as f:
        json.dump({"near_tie_eps": mod.NEAR_TIE_EPS, "mask_tie_eps": mod.MASK_TIE_EPS, "exitstatus": int(exitstatus), "stages": mod.RECORDS}, f, indent=0)
