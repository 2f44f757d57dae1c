import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionfinish(session, exitstatus):
    """Dump the worst errors the GPU parity tests measured (tests/common.py::record)."""
    import json
    import os
    import sys
    common = sys.modules.get("common")
    if common is None or not getattr(common, "MEASURED", None):
        return
    out = os.environ.get("DMB_PARITY_REPORT") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                              "gpurun_out", "parity_measured.json")
    try:
        os.makedirs(os.path.dirname(out), exist_ok=True)
        with open(out, "w") as f:
            json.dump(common.MEASURED, f, indent=1, sort_keys=True)
    except OSError:
        pass
