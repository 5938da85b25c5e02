import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def lib():
    from arshamg_scrnaseq_wgan_b200 import _lib
    _lib.build()
    return _lib.load()


def pytest_sessionfinish(session, exitstatus):
    """GPU runs: keep the measured gradient errors of the step tests next to their tolerances
    (gpurun_out/r02_step_errors.json; the tolerances in tests/test_steps_gpu.py are set from this file)."""
    mod = sys.modules.get("test_steps_gpu")
    log = getattr(mod, "MEASURED_LOG", None)
    if not log:
        return
    import json
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    worst = {}
    for tid, mode, B, k, fro, ftol, mx, mtol in log:
        key = f"{tid.split(' ')[0]}"
        w = worst.setdefault(key, {"mode": mode, "B": B, "fro": 0.0, "fro_tol": ftol, "max_rel": 0.0, "max_rel_tol": mtol})
        w["fro"], w["max_rel"] = max(w["fro"], fro), max(w["max_rel"], mx)
    with open(os.path.join(out, "r02_step_errors.json"), "w") as f:
        json.dump(worst, f, indent=1)
