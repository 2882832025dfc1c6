"""The reference loader lives in baseline/loader.py (bench.py's reference arm uses it too); tests keep importing `refshim`."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from baseline import loader as _loader  # noqa: E402

sys.modules[__name__] = _loader
