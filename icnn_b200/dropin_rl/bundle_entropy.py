"""Drop-in module named ``bundle_entropy`` for the RL call site (RL/src/icnn.py:8,155:
``bundle_entropy.solveBatch(fg, act)[0]``): signature and defaults of
RL/src/bundle_entropy.py:85 (nIter=5, clip to [0.03, 0.97], callback(t, f))."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from icnn_b200 import bundle_entropy as _be  # noqa: E402


def solveBatch(fg, initXs, nIter=5, callback=None, **kw):
    return _be.solveBatch(fg, initXs, nIter=nIter, callback=callback, variant='rl', **kw)
