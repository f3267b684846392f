"""TEST INFRASTRUCTURE ONLY -- inert stand-in (see ../matplotlib/__init__.py)."""
import sys as _sys
from unittest import mock as _mock

mplot3d = _mock.MagicMock(name="mpl_toolkits.mplot3d")
_sys.modules["mpl_toolkits.mplot3d"] = mplot3d
