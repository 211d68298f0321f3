"""Parity of the CUDA model builder (itr_build_model, through the C ABI) with the
reference's trans_emiss_calc outputs (tests/golden/model_*.npz, produced by running the
reference) and with the CPU oracle (oracle/ctmc_oracle.py) where the reference is
intractable.  Tolerances: a, pi 1e-12 relative; b 1e-8 relative + 2e-16 absolute (both
sides carry ~1e-17 cancellation noise in the coalescent integrals, entries go down to
4e-9) — the same bars tests/test_oracle_model.py holds the oracle to."""
import numpy as np
import pytest

import ctmc_oracle as co
from conftest import golden, golden_models

pytestmark = pytest.mark.gpu


def _check(a, b, pi, ra, rb, rpi, a_rtol=1e-12):
    np.testing.assert_allclose(a, ra, rtol=a_rtol, atol=1e-18)
    np.testing.assert_allclose(pi, rpi, rtol=a_rtol)
    np.testing.assert_allclose(b, rb, rtol=1e-8, atol=2e-16)
    assert abs(pi.sum() - 1) < 1e-12
    np.testing.assert_allclose(a.sum(1), 1, atol=1e-12)
    np.testing.assert_allclose(b.sum(1), 1, atol=1e-12)


@pytest.mark.parametrize("fn", golden_models())
def test_build_vs_reference_golden(engine, fn):
    g = golden(fn)
    n_ab, n_abc = (int(x) for x in g["n_int"])
    a, b, pi, hidden = engine.build_model(g["args"][None, :], n_ab, n_abc)
    assert np.array_equal(hidden, g["hidden"])
    _check(a[0], b[0], pi[0], g["a"], g["b"], g["pi"])


def _random_sets(n, seed):
    """Parameter sets drawn log-uniformly inside example_config.yaml's [min, max] boxes
    (BASELINE.md config 5), in the scaled units trans_emiss_calc takes."""
    rng = np.random.default_rng(seed)
    mu = 1e-8
    out = []
    for _ in range(n):
        lu = lambda lo, hi: float(np.exp(rng.uniform(np.log(lo), np.log(hi))))
        N_AB, N_ABC = lu(5e3, 5e5) * mu, lu(5e3, 5e5) * mu
        t_A, t_B = lu(24e3, 24e5) * mu, lu(24e3, 24e5) * mu
        t_2 = lu(4e3, 4e5) * mu
        t_C = (t_A + t_B) / 2 + t_2
        t_upper = lu(74e3, 74e5) * mu
        r = lu(1e-9, 1e-7) / mu
        t_out = t_C + 3 * N_ABC + t_upper
        out.append([t_A, t_B, t_C, t_2, t_upper, t_out, N_AB, N_ABC, r])
    return np.array(out)


@pytest.mark.parametrize("n_ab,n_abc", [(3, 3), (2, 4), (1, 1)])
def test_batched_build_vs_oracle(engine, n_ab, n_abc):
    params = _random_sets(6, 20261018 + n_ab * 10 + n_abc)
    a, b, pi, hidden = engine.build_model(params, n_ab, n_abc)
    for s in range(len(params)):
        ra, rb, rpi, hid, _ = co.trans_emiss_calc(*params[s], n_ab, n_abc)
        assert [tuple(h) for h in hidden] == [hid[i] for i in range(len(hid))]
        # extreme corners of the box put ~1e-11 relative noise on the smallest
        # transition probabilities of both implementations
        _check(a[s], b[s], pi[s], ra, rb, rpi, a_rtol=1e-10)


def test_finer_discretisation_vs_oracle(engine):
    g = golden("model_3_3_example.npz")
    a, b, pi, hidden = engine.build_model(g["args"][None, :], 5, 5)
    ra, rb, rpi, hid, _ = co.trans_emiss_calc(*g["args"], 5, 5)
    assert a.shape == (1, 70, 70)
    _check(a[0], b[0], pi[0], ra, rb, rpi)


def test_custom_cutpoints(engine):
    g = golden("model_2_2_example.npz")
    args = g["args"]
    N_ref = args[7]
    t_AB, coal_AB = args[3] / N_ref, N_ref / args[6]
    cut_AB = np.array([0.0, 0.3 * t_AB, t_AB])
    cut_ABC = np.array([0.0, 0.5, np.inf])
    a, b, pi, _ = engine.build_model(args[None, :], 2, 2, cut_AB, cut_ABC)
    ra, rb, rpi, _, _ = co.trans_emiss_calc(*args, 2, 2, cut_AB, cut_ABC)
    _check(a[0], b[0], pi[0], ra, rb, rpi)


def test_trans_emiss_calc_signature_and_loglik(engine):
    """The reference-style function returns the reference's five-tuple, and the model
    left on the device by the builder gives the same log-likelihood as installing the
    fetched matrices."""
    import itrails_b200 as itb
    import hmm_oracle as ho
    g = golden("model_3_3_example.npz")
    a, b, pi, hidden_names, observed_names = itb.trans_emiss_calc(*g["args"], 3, 3)
    assert hidden_names[0] == (0, 0, 0) and hidden_names[26] == (3, 2, 2)
    assert observed_names[0] == "AAAA" and observed_names[255] == "GGGG" and observed_names[2] == "AAAT"
    np.testing.assert_allclose(a, g["a"], rtol=1e-12, atol=1e-18)
    rng = np.random.default_rng(5)
    V_lst = [ho.sample_block(a, b, pi, T, rng, p_n=0.01) for T in (5000, 1234)]
    engine.load_blocks(V_lst)
    engine.build_model(g["args"][None, :], 3, 3, fetch=False)
    ll_dev = engine.loglik()[0]
    engine.set_model(a, b, pi)
    ll_host = engine.loglik()[0]
    assert ll_dev == ll_host
    ref = float(ho.loglik_wrapper(g["a"], g["b"], g["pi"], V_lst))
    assert abs(ll_dev - ref) <= 1e-9 * abs(ref)


def test_bad_parameters_are_rejected(engine):
    g = golden("model_1_1_example.npz")
    bad = g["args"].copy()
    bad[6] = -1.0
    with pytest.raises(ValueError):
        engine.build_model(bad[None, :], 1, 1)
