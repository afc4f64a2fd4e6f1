"""Locate and import the cokrig_b200 binding from the drop-in modules.

The drop-in modules are imported the way the reference's notebooks import theirs
(``sys.path.insert(0, ".../src")``; research/simulation_experiment.ipynb[1]), so the binding
package one directory up is put on ``sys.path`` here.  Import failure (library not built) is NOT
swallowed: there is no CPU fallback.
"""
import os
import sys

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG_ROOT not in sys.path:
    sys.path.insert(0, _PKG_ROOT)

import cokrig_b200  # noqa: E402
from cokrig_b200 import ops  # noqa: E402,F401
