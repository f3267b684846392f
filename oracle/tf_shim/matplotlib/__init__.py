"""TEST INFRASTRUCTURE ONLY -- inert stand-in for matplotlib so the reference's unit-test modules (which
import plotting helpers at module level) can be imported by oracle/run_reference_unittests.py."""
import sys as _sys
import types as _types
from unittest import mock as _mock

for _name in ("axes", "pyplot", "cm", "colors", "patches", "gridspec", "ticker", "animation", "backends",
              "backends.backend_agg", "figure", "lines", "collections"):
    _m = _mock.MagicMock(name="matplotlib." + _name)
    _sys.modules["matplotlib." + _name] = _m
    globals()[_name.split(".")[0]] = _m


def use(*a, **k):
    pass


rcParams = {}
