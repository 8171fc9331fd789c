import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _parity_case(request):
    """Name the parity-report case after the running gpu test (tests/util.py records under it)."""
    import util
    if "gpu" in request.keywords:
        util.CURRENT_CASE[0] = request.node.name[5:] if request.node.name.startswith("test_") else request.node.name
        if request.node.fspath.basename == "test_kernels_gpu.py":
            util.KERNEL_CASES.add(util.CURRENT_CASE[0].split("[")[0])
    yield
    util.CURRENT_CASE[0] = None


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    """Parity audit: one line per recorded parity case (shown even under -q, so it lands in the driver's GPUTEST
    tail) and the per-parameter table in profiles/parity_report.json (+ gpurun_out/ so it travels back)."""
    import util
    if not util.REPORT and not util.NOTES:
        return
    tr = terminalreporter
    tr.section("parity report (tests/util.py)")
    for line in util.summarize():
        tr.write_line(line)
    for note in util.NOTES:
        tr.write_line(note)
    doc = {"tolerances": {"fp32": 1e-3, "bf16": 2e-2}, "exemptible_parameters": list(util.EXEMPTIBLE),
           "exemption_rule": "bf16 only: err may exceed tol if stock torch.autocast(bf16) of the fp32 oracle also does, "
                             "and then by <= 1.5x that yardstick error",
           "summary": util.summarize(), "notes": util.NOTES, "entries": util.REPORT}
    for d in ("profiles", "gpurun_out"):
        try:
            os.makedirs(os.path.join(ROOT, d), exist_ok=True)
            with open(os.path.join(ROOT, d, "parity_report.json"), "w") as f:
                json.dump(doc, f, indent=1)
        except OSError:
            pass
