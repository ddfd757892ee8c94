"""Drop-in module named ``bundle_entropy`` for the reference's lib/ call sites.

completion/icnn_ebundle.py:28-31 and multi-label-cls/icnn_ebundle.py:27-30 do
``sys.path.append('../lib'); import bundle_entropy``; pointing that path entry at this directory
instead makes ``bundle_entropy.solveBatch(fg, y0, nIter=..., callback=..., solver='pc')`` run on
the GPU with the same signature and return tuple (lib/bundle_entropy.py:192,242)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from icnn_b200 import bundle_entropy as _be  # noqa: E402


def solveBatch(fg, initXs, nIter=10, callback=None, solver='pc', **kw):
    return _be.solveBatch(fg, initXs, nIter=nIter, callback=callback, solver=solver, variant='lib', **kw)


def solve(fg, initX, nIter=10, callback=None, **kw):
    return _be.solve(fg, initX, nIter=nIter, callback=callback, variant='lib', **kw)
