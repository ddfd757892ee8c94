"""icnn_b200 -- B200-native inner loop argmin_y f(x, y; theta) of locuslab/icnn.

Public surface (mirrors the reference's):
    icnn_b200.bundle_entropy.solveBatch(fg, initXs, nIter, callback, solver, variant=...)
    icnn_b200.PICNN(...).bind(x)         -> the fg object for the fused on-device loop
    icnn_b200.gd.solve(...)              -> unrolled momentum gradient descent
    icnn_b200.argmin_grad.argmin_grad(state, trueY, loss) -> crossEntrGrad / mseGrad + train_step_fd feeds
    icnn_b200.gd_grad.gd_grad(fg, y0, trueY, ...) -> d mse / d theta through the unrolled GD loop
The compute path is hand-written sm_100a CUDA behind a C ABI (libicnn_b200.so); importing this
package without the built library raises ImportError -- there is no CPU fallback.
"""
from . import _capi  # noqa: F401  (fails loudly if the native library is missing)
from .picnn import PICNN, BoundPICNN  # noqa: F401
from . import bundle_entropy, gd, argmin_grad, adam, gd_grad  # noqa: F401

__all__ = ["PICNN", "BoundPICNN", "bundle_entropy", "gd", "argmin_grad", "adam", "gd_grad"]
