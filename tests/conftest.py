import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


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


# ---------------------------------------------------------------------------------------------------------------
# B200DIFF_PARITY_LOG=<path>: write every parity quantity the suite computes (O.rel_err / _util.maxabs: value, test id,
# file:line of the assert that used it) to <path> as JSON.  This is how the bounds in tests/test_gpu_*.py are derived
# ("measured x margin", profiles/r2_parity_measured.md) instead of being round numbers.
# ---------------------------------------------------------------------------------------------------------------
_PARITY_LOG = os.environ.get("B200DIFF_PARITY_LOG")
_records, _current = [], [""]

if _PARITY_LOG:
    import dsdiff_oracle as _O
    import _util as _U

    def _recording(fn, name):
        def wrapped(a, b):
            v = fn(a, b)
            fr = sys._getframe(1)
            _records.append({"test": _current[0], "file": os.path.basename(fr.f_code.co_filename), "line": fr.f_lineno,
                             "fn": name, "value": float(v)})
            return v
        wrapped._recording = True
        return wrapped

    _O.rel_err = _recording(_O.rel_err, "rel_err")
    _U.maxabs = _recording(_U.maxabs, "maxabs")

    def pytest_itemcollected(item):
        mod = getattr(item, "module", None)          # `from _util import maxabs` bound the plain function at import
        fn = getattr(mod, "maxabs", None)
        if fn is not None and not getattr(fn, "_recording", False):
            mod.maxabs = _U.maxabs

    def pytest_runtest_setup(item):
        _current[0] = item.nodeid

    def pytest_sessionfinish(session, exitstatus):
        import json
        with open(_PARITY_LOG, "w") as f:
            json.dump(_records, f, indent=0)
