"""Property tests (hypothesis) of the invariants the reference code implies:
for any bundle (G, h) the subproblem solvers agree on the unique optimum, the multipliers live on
the simplex, y = sigma(-G^T lam), and primal = dual value.  CPU only."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import bundle_np


@st.composite
def bundles(draw):
    k = draw(st.integers(1, 5))
    n = draw(st.integers(k, 12))
    seed = draw(st.integers(0, 2**31 - 1))
    rs = np.random.RandomState(seed)
    scale = draw(st.sampled_from([0.3, 1.0, 3.0]))
    return rs.randn(k, n) * scale, rs.randn(k) * scale


@settings(max_examples=40, deadline=None)
@given(bundles())
def test_subproblem_invariants(gh):
    G, h = gh
    k, n = G.shape
    with np.errstate(all="ignore"):
        y, z = bundle_np.pdipm_pc(G, h)
        lam = bundle_np.proj_newton_logistic(G, h, line_search=True)
    yd = 1.0 / (1.0 + np.exp(G.T.dot(lam)))
    assert np.all(lam >= 0) and abs(lam.sum() - 1.0) < 1e-9                 # simplex
    assert np.all(z > 0) and abs(z.sum() - 1.0) < 1e-6
    assert np.abs(y - yd).max() < 1e-5                                       # PC and dual agree
    assert np.all((y > 0) & (y < 1))
    primal = np.max(G.dot(yd) + h) + bundle_np.neg_entropy(yd)
    dual = (G.sum(axis=1) + h).dot(lam) - np.sum(bundle_np.softplus(G.T.dot(lam)))
    assert abs(primal - dual) < 1e-6 * max(1.0, abs(primal))                 # strong duality
    if k == 1:
        np.testing.assert_allclose(yd, 1.0 / (1.0 + np.exp(G[0])), atol=1e-12)   # one cut: y = sigma(-g)


def test_reference_cost_mode_is_the_same_algorithm():
    """bench.py's `cpu_baseline.reference_cost` runs the port with the reference's dense np.diag matrices and
    per-iteration prints (lib/bundle_entropy.py:17-18,34-36,41): same numbers, only slower."""
    import contextlib
    import io
    from oracle import bundle_np, picnn_np, synth
    p, x, y0 = synth.make_inputs("C1", B=6)
    fg = picnn_np.make_fg(p, x)
    with np.errstate(all="ignore"):
        a = bundle_np.solve_batch(fg, y0.copy(), nIter=4)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            b = bundle_np.solve_batch(fg, y0.copy(), nIter=4, dense_diag=True, verbose=True)
    np.testing.assert_allclose(a[0], b[0], atol=1e-10)
    assert "primal_res" in buf.getvalue() and "kappa(d)" in buf.getvalue()
