"""Shadow module: put this directory first on sys.path and the reference's own
`from VBx import VBx` (VBx/vbhmm.py:45) resolves to the B200 implementation, unchanged call site."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from vbx_b200.api import VBx, DER  # noqa: E402,F401
