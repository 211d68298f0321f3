"""Lock-step FP64 tensor-core sweeps (csrc/lockstep.cu, 32 < K <= 96, many blocks) against
the C oracle: forward log-likelihood 1e-9 relative, posterior 1e-7 absolute
(BASELINE.json's tolerances).  Every template size (K rounded up to 8: 40 ... 96), ragged
block lengths incl. 1, 2, 3 (odd lengths delay the backward row by a step), block counts
that are not multiples of the group size, several parameter sets in one call, and
agreement with the one-CTA-per-chain sweeps the library uses when blocks are few."""
import numpy as np
import pytest

import hmm_oracle as ho
import hmm_oracle_c as hoc

pytestmark = pytest.mark.gpu

# (n_int_AB, n_int_ABC) -> K = n_AB n_ABC + 3 n_ABC + 3 C(n_ABC, 2)
DISCRETISATIONS = [(5, 3), (3, 4), (5, 4), (3, 5), (5, 5), (6, 5), (3, 6), (5, 6)]
LENS = [1, 2, 3, 4, 5, 7, 8, 9, 16, 17, 31, 32, 33, 64, 100, 255, 257, 700, 1001, 1500, 2, 9, 333]


def _model(engine, n_ab, n_abc):
    from itrails_b200 import synth
    a, b, pi, _ = engine.build_model(synth.example_model_args(n_abc)[None, :], n_ab, n_abc)
    return a[0], b[0], pi[0]


def _blocks(a, b, pi, lens, seed):
    rng = np.random.default_rng(seed)
    V_lst = [ho.sample_block(a, b, pi, T, rng, p_n=0.03) for T in lens]
    V_lst[10][:] = 624             # all-N block
    return V_lst


@pytest.mark.parametrize("n_ab,n_abc", DISCRETISATIONS)
def test_lockstep_vs_oracle(engine, monkeypatch, n_ab, n_abc):
    a, b, pi = _model(engine, n_ab, n_abc)
    K = a.shape[0]
    assert 32 < K <= 96
    V_lst = _blocks(a, b, pi, LENS, 20261018 + K)
    E = ho.emission_table(b)
    ref_ll = hoc.loglik_blocks(a, E, pi, V_lst)
    ref_post = hoc.post_prob_blocks(a, E, pi, V_lst)
    engine.load_blocks(V_lst)
    engine.set_model(a, b, pi)
    monkeypatch.setenv("ITR_LOCKSTEP", "1")
    n0 = engine.lockstep_launch_count
    tot, pb = engine.loglik(per_block=True)
    np.testing.assert_allclose(pb[0], ref_ll, rtol=1e-9, atol=1e-12)
    assert abs(tot[0] - ref_ll.sum()) <= 1e-9 * abs(ref_ll.sum())
    post = engine.split(engine.posterior())
    assert engine.lockstep_launch_count - n0 == 2          # both calls took the lock-step kernels
    for p, r in zip(post, ref_post):
        assert p.shape == r.shape
        assert np.abs(p - r).max() <= 1e-7
        np.testing.assert_allclose(p.sum(1), 1.0, atol=1e-12)
    # the sweeps used for few blocks agree (both are within tolerance of the oracle; this pins
    # the two code paths to each other more tightly)
    monkeypatch.setenv("ITR_LOCKSTEP", "0")
    tot2, pb2 = engine.loglik(per_block=True)
    np.testing.assert_allclose(pb2[0], pb[0], rtol=1e-12, atol=1e-13)
    post2 = engine.split(engine.posterior())
    for p, q in zip(post, post2):
        assert np.abs(p - q).max() <= 1e-10


def test_lockstep_many_sets(engine, monkeypatch):
    """Three parameter sets x 13 blocks in one call: groups never mix sets."""
    from itrails_b200 import synth
    base = synth.example_model_args(5)
    params = np.stack([base, base * np.array([1.1, 1.1, 1.05, 0.9, 1.0, 1.02, 1.2, 0.9, 1.3]),
                       base * np.array([0.8, 0.8, 0.9, 1.2, 1.1, 0.97, 0.7, 1.15, 0.6])])
    a, b, pi, _ = engine.build_model(params, 5, 5)
    lens = [900, 5, 64, 1200, 33, 1, 2, 777, 31, 450, 8, 100, 2000]
    V_lst = _blocks(a[0], b[0], pi[0], lens, 99)
    engine.load_blocks(V_lst)
    engine.set_model(a, b, pi)
    monkeypatch.setenv("ITR_LOCKSTEP", "1")
    tot, pb = engine.loglik(per_block=True)
    for s in range(3):
        ref = hoc.loglik_blocks(a[s], ho.emission_table(b[s]), pi[s], V_lst)
        np.testing.assert_allclose(pb[s], ref, rtol=1e-9, atol=1e-12)
        assert abs(tot[s] - ref.sum()) <= 1e-9 * abs(ref.sum())


def test_lockstep_is_the_default_for_many_blocks(engine, monkeypatch):
    """600 short blocks at K = 70: the library picks the lock-step kernels by itself, copies of
    a block give bit-identical rows whatever group, row and warp rotation they land in."""
    monkeypatch.delenv("ITR_LOCKSTEP", raising=False)
    a, b, pi = _model(engine, 5, 5)
    base = _blocks(a, b, pi, [300, 301, 64, 65, 1, 17, 500, 299, 2, 1000, 77, 12], 5)
    V_lst = base * 50
    engine.load_blocks(V_lst)
    engine.set_model(a, b, pi)
    n0 = engine.lockstep_launch_count
    tot, pb = engine.loglik(per_block=True)
    post = engine.split(engine.posterior())
    assert engine.lockstep_launch_count - n0 == 2
    pb = pb[0].reshape(50, len(base))
    assert np.array_equal(pb, np.broadcast_to(pb[0], pb.shape))
    E = ho.emission_table(b)
    np.testing.assert_allclose(pb[0], hoc.loglik_blocks(a, E, pi, base), rtol=1e-9, atol=1e-12)
    ref_post = hoc.post_prob_blocks(a, E, pi, base)
    for i, r in enumerate(ref_post):
        assert np.abs(post[i] - r).max() <= 1e-7
        for c in (1, 23, 49):
            assert np.array_equal(post[i + c * len(base)], post[i]), (i, c)
