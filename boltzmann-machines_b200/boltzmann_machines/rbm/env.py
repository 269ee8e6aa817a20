"""The reference's tests import ``from rbm import ...`` / ``from utils import ...``
after this module prepends the package directory to ``sys.path``
(/root/reference/boltzmann_machines/rbm/env.py).  Kept so such imports resolve."""
import os.path as _p
import sys as _sys

_pkg_dir = _p.dirname(_p.dirname(_p.abspath(__file__)))
if _pkg_dir not in _sys.path:
    _sys.path.insert(0, _pkg_dir)
