import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with `pytest -m gpu`")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a GPU-less box would otherwise fail every test on sb_create; skip with the reason instead
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (run through gpurun)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


# ---- parity report: every disagreement the GPU parity tests see (keypoints unmatched either way, descriptor rows over
# 1e-4, match rows that differ) is collected here and written at the end of a GPU session to gpurun_out/ (which travels
# back from the GPU box) and to profiles/parity_report.json (committed copy).
class _ParityReport:
    def __init__(self):
        self.entries = {}

    def add(self, test, **kw):
        self.entries.setdefault(test, []).append(kw)


PARITY = _ParityReport()


@pytest.fixture
def report(request):
    name = request.node.name

    def add(**kw):
        PARITY.add(name, **kw)
    return add


def pytest_sessionfinish(session, exitstatus):
    if not PARITY.entries:
        return
    import json
    doc = {"what": "disagreements seen by `pytest -m gpu` (tests/test_parity_gpu.py, tests/test_demo_gpu.py); empty lists = none",
           "exitstatus": int(exitstatus), "tests": PARITY.entries}
    for d in ("gpurun_out", "profiles"):
        try:
            os.makedirs(os.path.join(ROOT, d), exist_ok=True)
            with open(os.path.join(ROOT, d, "parity_report.json"), "w") as f:
                json.dump(doc, f, indent=1, default=float)
        except OSError:
            pass
