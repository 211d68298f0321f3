"""Pins the recursion oracle (NumPy and C restatements) against fixtures produced by
running the reference itself (tests/golden/make_golden.py)."""
import numpy as np
import pytest

import hmm_oracle as ho
import hmm_oracle_c as hoc
from conftest import golden, golden_recursions


def test_symbol_alphabet_matches_reference():
    g = golden("symbols.npz")
    assert list(g["names"]) == ho.obs_state_names()
    off, vals = g["order_offsets"], g["order_values"]
    order = ho.order_lists()
    assert len(order) == 625
    for s in range(625):
        assert np.array_equal(order[s], vals[off[s]:off[s + 1]])
    assert list(order[256]) == [0, 1, 2, 3] and len(order[624]) == 256


@pytest.mark.parametrize("fn", golden_recursions())
def test_numpy_oracle_vs_reference(fn):
    g = golden(fn)
    m = golden(str(g["model_file"]))
    a, b, pi = m["a"], m["b"], m["pi"]
    E = ho.emission_table(b)
    tot = 0.0
    for i in range(int(g["n_blocks"])):
        V = g[f"V_{i}"]
        alpha = ho.forward(a, b, pi, V, E)
        np.testing.assert_allclose(alpha, g[f"alpha_{i}"], rtol=1e-12, atol=1e-11)
        ll = ho.forward_loglik(a, b, pi, V, E)
        assert abs(ll - g[f"loglik_{i}"]) <= 1e-12 * abs(g[f"loglik_{i}"]) + 1e-12
        tot += ll
        np.testing.assert_allclose(ho.backward(a, b, V, E), g[f"beta_{i}"], rtol=1e-12, atol=1e-11)
        assert np.abs(ho.post_prob(a, b, pi, V, E) - g[f"post_{i}"]).max() < 1e-12
        if len(V) > 1:
            om, prev = ho.viterbi(a, b, pi, V, E)
            assert np.array_equal(ho.backtrack_viterbi(om, prev), g[f"vit_{i}"])      # bit-exact
            assert np.array_equal(om[-1], g[f"vit_omega_last_{i}"])                   # bit-exact
    assert abs(tot - g["loglik_total"]) <= 1e-12 * abs(g["loglik_total"])


@pytest.mark.parametrize("fn", golden_recursions())
def test_c_oracle_vs_reference(fn):
    g = golden(fn)
    m = golden(str(g["model_file"]))
    a, b, pi = m["a"], m["b"], m["pi"]
    n = int(g["n_blocks"])
    V_lst = [g[f"V_{i}"] for i in range(n)]
    E = ho.emission_table(b)
    for threads in (1, 3):
        ll = hoc.loglik_blocks(a, E, pi, V_lst, threads)
        for i in range(n):
            assert abs(ll[i] - g[f"loglik_{i}"]) <= 1e-12 * abs(g[f"loglik_{i}"]) + 1e-12
        post = hoc.post_prob_blocks(a, E, pi, V_lst, threads)
        for i in range(n):
            assert np.abs(post[i] - g[f"post_{i}"]).max() < 1e-12
        LA, LE, om0 = ho.viterbi_tables(a, b, pi, V_lst)
        vit = hoc.viterbi_blocks(LA, LE, om0, V_lst, threads)
        for i in range(n):
            assert np.array_equal(vit[i].astype(np.float64), g[f"vit_{i}"])           # bit-exact


def test_backward_orientation_is_the_references():
    """The reference multiplies the row vector by a (optimizer.py:210); the textbook
    recursion (a @ v) gives visibly different posteriors — guard against 'fixing' it."""
    g = golden(golden_recursions()[0])
    m = golden(str(g["model_file"]))
    a, b, pi = m["a"], m["b"], m["pi"]
    V = g["V_0"][:300]
    E = ho.emission_table(b)
    ref = ho.post_prob(a, b, pi, V, E)
    T, K = len(V), a.shape[0]
    beta = np.zeros((T, K))
    for t in range(T - 2, -1, -1):
        x = beta[t + 1].max()
        beta[t] = np.log(a @ (np.exp(beta[t + 1] - x) * E[:, V[t + 1]])) + x
    s = ho.forward(a, b, pi, V, E) + beta
    w = np.exp(s - s.max(1, keepdims=True))
    textbook = w / w.sum(1, keepdims=True)
    assert np.abs(textbook - ref).max() > 1e-4


def test_sample_block_is_deterministic_and_in_range():
    m = golden("model_2_2_example.npz")
    V1 = ho.sample_block(m["a"], m["b"], m["pi"], 5000, np.random.default_rng(3), p_n=0.05)
    V2 = ho.sample_block(m["a"], m["b"], m["pi"], 5000, np.random.default_rng(3), p_n=0.05)
    assert np.array_equal(V1, V2) and V1.min() >= 0 and V1.max() <= 624 and (V1 >= 256).any()
