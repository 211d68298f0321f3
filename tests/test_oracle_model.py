"""Pins the model-build oracle (oracle/ctmc_oracle.py) against (a, b, pi) produced by
the reference's trans_emiss_calc (tests/golden/model_*.npz)."""
import numpy as np
import pytest
import scipy.linalg

import ctmc_oracle as co
from conftest import golden, golden_models


@pytest.mark.parametrize("fn", golden_models())
def test_model_vs_reference(fn):
    g = golden(fn)
    n_ab, n_abc = (int(x) for x in g["n_int"])
    a, b, pi, hid, obs = co.trans_emiss_calc(*g["args"], n_ab, n_abc)
    assert np.array_equal(np.array([hid[i] for i in range(len(hid))]), g["hidden"])
    assert [obs[i] for i in range(256)] == list(g["observed"])
    # transition rows / start vector: 1e-12 relative
    np.testing.assert_allclose(a, g["a"], rtol=1e-12, atol=1e-18)
    np.testing.assert_allclose(pi, g["pi"], rtol=1e-12)
    # emissions: both sides carry ~1e-17 absolute rounding noise from cancellation in
    # the coalescent integrals (entries go down to ~4e-9), hence absolute + relative.
    np.testing.assert_allclose(b, g["b"], rtol=1e-8, atol=2e-16)
    # invariants the reference's output satisfies
    assert abs(pi.sum() - 1) < 1e-12
    np.testing.assert_allclose(a.sum(1), 1, atol=1e-12)
    np.testing.assert_allclose(b.sum(1), 1, atol=1e-12)
    J = a * pi[:, None]
    assert np.abs(J - J.T).max() < 1e-15


def test_state_spaces_match_reference():
    g = golden("statespace.npz")
    for n, size, ntrans in ((1, 2, 2), (2, 15, 44), (3, 203, 1118)):
        ss = co.state_space(n)
        assert ss.size == size and len(ss.trans) == ntrans
        ref_states = [tuple(r) for r in g[f"states_{n}"]]
        assert sorted(ref_states) == sorted(ss.states)
        tr = g[f"transitions_{n}"]
        m = 2 * n
        ref_tr = sorted((tuple(r[:m]), tuple(r[m:2 * m]), int(r[2 * m + 2])) for r in tr)
        mine = sorted((ss.states[f], ss.states[t], int(k)) for f, t, k in ss.trans)
        assert ref_tr == mine
        # omega classes
        ref_map = {}
        for key, mask in zip(g[f"omega_keys_{n}"], g[f"omega_masks_{n}"]):
            for i in np.nonzero(mask)[0]:
                ref_map[ref_states[i]] = tuple(int(x) for x in key)
        for s, om in zip(ss.states, ss.omega):
            assert ref_map[s] == om


def test_expm_against_scipy():
    rng = np.random.default_rng(0)
    ss = co.state_space(3)
    Q = ss.generator(1.0, 0.37)
    for t in (1e-3, 0.4, 3.0, 40.0):
        np.testing.assert_allclose(co.expm(Q * t), scipy.linalg.expm(Q * t), atol=2e-14)
    A = rng.normal(size=(12, 12))
    np.testing.assert_allclose(co.expm(A), scipy.linalg.expm(A), rtol=1e-12, atol=1e-12)


def test_cutpoints_against_scipy():
    from scipy.stats import expon, truncexpon
    for n, t, c in ((3, 0.8, 1.0), (5, 2.5, 0.4), (1, 0.1, 3.0)):
        q = np.arange(n + 1) / n
        np.testing.assert_allclose(co.cutpoints_AB(n, t, c), truncexpon.ppf(q, b=t * c, scale=1 / c), rtol=1e-14, atol=1e-16)
        np.testing.assert_allclose(co.cutpoints_ABC(n, c), expon.ppf(q, scale=1 / c), rtol=1e-14)


def test_larger_discretisation_invariants():
    """(3,4): the reference needs ~an hour here; the oracle must still satisfy the
    reference's invariants (parity at this size is transitive)."""
    g = golden("model_1_1_example.npz")
    a, b, pi, hid, _ = co.trans_emiss_calc(*g["args"], 3, 4)
    K = 3 * 4 + 3 * 4 + 3 * 6
    assert a.shape == (K, K) and b.shape == (K, 256)
    assert abs(pi.sum() - 1) < 1e-12 and np.abs(a.sum(1) - 1).max() < 1e-12
    J = a * pi[:, None]
    assert np.abs(J - J.T).max() < 1e-15 and (a > 0).all()
    # t_A == t_B  =>  topologies 2 and 3 are exchangeable
    i2 = [k for k, h in hid.items() if h[0] == 2]
    i3 = [k for k, h in hid.items() if h[0] == 3]
    np.testing.assert_allclose(pi[i2], pi[i3], rtol=1e-12)
