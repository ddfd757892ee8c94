"""RL Adam argmin (RL/src/icnn.py:160-215) on the GPU vs the numpy restatement and the golden vector
produced by the reference's own method body (oracle/gen_golden_adam.py)."""
import os

import numpy as np
import pytest

from oracle import adam_np, synth

pytestmark = pytest.mark.gpu


def test_adam_matches_reference_golden(golden_dir):
    import icnn_b200
    gold = np.load(os.path.join(golden_dir, "adam.npz"))
    p, x, _ = synth.make_inputs("C4", B=96)
    net = icnn_b200.PICNN.from_params(p)
    act, its = icnn_b200.adam.solve(net.bind(x), return_iters=True)
    assert act.shape == (96, 6) and np.abs(act).max() <= 1.0
    # float32 negQ / gradient under 78 Adam steps: the stop iteration may move by a few steps
    assert abs(its - int(gold["c4_iters"])) <= 5
    d = np.abs(act - gold["c4_act_best"]).max(axis=1)
    assert np.median(d) < 1e-4 and np.mean(d < 5e-3) >= 0.9, (np.median(d), d.max())


@pytest.mark.parametrize("B", [7, 300])
def test_adam_matches_oracle_objective(B):
    """Same optimiser, float32 f/g: the best objective found matches the oracle's."""
    import icnn_b200
    p, x, _ = synth.make_inputs("C4", B=B, seed=11)
    func = adam_np.make_fg_entr(p, x)
    obest, oits = adam_np.adam(func, x, p.n)
    net = icnn_b200.PICNN.from_params(p)
    act, its = icnn_b200.adam.solve(net.bind(x), return_iters=True)
    fo = func(x, obest)[0]
    fgpu = func(x, act)[0]
    assert abs(its - oits) <= 5
    assert np.median(np.abs(fgpu - fo)) < 1e-4 * max(1.0, np.abs(fo).max())
    assert np.all(fgpu <= fo + 1e-2 * max(1.0, np.abs(fo).max()))


def test_adam_rejects_affine_binding():
    import icnn_b200
    p, x, _ = synth.make_inputs("C4", B=4)
    net = icnn_b200.PICNN.from_params(p)
    with pytest.raises(ValueError):
        icnn_b200.adam.solve(net.bind(x, affine=True))
