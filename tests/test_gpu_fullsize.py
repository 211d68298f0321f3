"""BASELINE.json's full-size configurations on one B200, checked through size-independent
properties (the oracle cannot run 250 Mb in a test):

* the alignment is a seeded 10 Mb set of blocks repeated n times, so block i + c*n_base is a
  copy of block i: per-block log-likelihoods, Viterbi paths and posterior matrices of a
  copy must be bit-identical to the first occurrence, however far into the 64-bit
  address range the copy lives;
* posterior rows sum to one; the total log-likelihood is the sum of the per-block values;
* the shortest blocks are checked against the C oracle directly (tolerances of BASELINE.json).

Results stay on the device and are read block by block (itr_*_fetch_range), as the CSV
writers do."""
import numpy as np
import pytest

import hmm_oracle as ho
import hmm_oracle_c as hoc
from conftest import golden

pytestmark = pytest.mark.gpu


def _tiled(V_base, copies):
    lens = np.array([len(v) for v in V_base] * copies, dtype=np.int64)
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    sym = np.tile(np.concatenate(V_base).astype(np.uint16), copies)
    return sym, off


def test_config4_250mb_viterbi_posterior_loglik():
    """Config 4: 2 500 blocks, 250 Mb, K = 27 on one GPU."""
    import itrails_b200 as itb
    from itrails_b200 import synth
    from itrails_b200.optimizer import viterbi_tables
    m = golden("model_3_3_example.npz")
    a, b, pi = m["a"], m["b"], m["pi"]
    rng = np.random.default_rng(20261018 + 4)
    lens = synth.block_lengths(100, 10_000_000, rng)
    lens[7] = 3_000                      # short blocks the oracle can check directly
    lens[63] = 5_001
    V_base = synth.alignment(a, b, pi, lens, 20261018 + 4)
    copies, nb = 25, len(V_base)
    sym, off = _tiled(V_base, copies)
    assert off[-1] > 2**31 // 27 * 3      # the posterior index range needs 64 bits
    with itb.Engine(0) as eng:
        eng.load_packed(sym, off)
        eng.set_model(a, b, pi)
        tot, pb = eng.loglik(per_block=True)
        pb = pb[0].reshape(copies, nb)
        assert np.array_equal(pb, np.broadcast_to(pb[0], pb.shape))
        assert abs(tot[0] - pb.sum()) <= 1e-12 * abs(tot[0])
        E = ho.emission_table(b)
        ref = hoc.loglik_blocks(a, E, pi, [V_base[7], V_base[63]])
        np.testing.assert_allclose(pb[0, [7, 63]], ref, rtol=1e-9)

        LA, LE, om0 = viterbi_tables(a, b, pi, V_base)
        eng.viterbi(LA, LE, np.tile(om0, (copies, 1)), fetch=False)
        eng.posterior(fetch=False)
        ref_path = hoc.viterbi_blocks(LA, LE, om0[[7, 63]], [V_base[7], V_base[63]])
        ref_post = hoc.post_prob_blocks(a, E, pi, [V_base[7], V_base[63]])
        for k, i in enumerate((7, 63)):
            assert np.array_equal(eng.viterbi_block(i), ref_path[k])
            assert np.abs(eng.posterior_block(i) - ref_post[k]).max() <= 1e-7
        for i in (0, 7, 41, 63, 99):       # first copy vs middle and last copies
            p0, q0 = eng.viterbi_block(i), eng.posterior_block(i)
            np.testing.assert_allclose(q0.sum(1), 1.0, atol=1e-12)
            for c in (11, copies - 1):
                assert np.array_equal(eng.viterbi_block(i + c * nb), p0), (i, c)
                assert np.array_equal(eng.posterior_block(i + c * nb), q0), (i, c)
        with pytest.raises(ValueError):
            eng._ck(eng._lib.itr_posterior_fetch_range(eng._ctx, int(off[-1]) - 1, 2, None))


def test_config3_100mb_posterior_finer_discretisation():
    """Config 3: 100 Mb at (n_int_AB, n_int_ABC) = (5, 5), K = 70, posterior kept in HBM."""
    import itrails_b200 as itb
    from itrails_b200 import synth
    rng = np.random.default_rng(20261018 + 3)
    with itb.Engine(0) as eng:
        args = synth.example_model_args(5)
        a, b, pi, _ = eng.build_model(args[None, :], 5, 5)
        a, b, pi = a[0], b[0], pi[0]
        assert a.shape == (70, 70)
        lens = synth.block_lengths(100, 10_000_000, rng)
        lens[5] = 4_000
        V_base = synth.alignment(a, b, pi, lens, 20261018 + 3)
        copies, nb = 10, len(V_base)
        sym, off = _tiled(V_base, copies)
        eng.load_packed(sym, off)
        tot, pb = eng.loglik(per_block=True)
        pb = pb[0].reshape(copies, nb)
        assert np.array_equal(pb, np.broadcast_to(pb[0], pb.shape))
        E = ho.emission_table(b)
        np.testing.assert_allclose(pb[0, 5], hoc.loglik_blocks(a, E, pi, [V_base[5]])[0], rtol=1e-9)
        eng.posterior(fetch=False)
        ref = hoc.post_prob_blocks(a, E, pi, [V_base[5]])[0]
        assert np.abs(eng.posterior_block(5) - ref).max() <= 1e-7
        for i in (0, 5, 99):
            q0 = eng.posterior_block(i)
            np.testing.assert_allclose(q0.sum(1), 1.0, atol=1e-12)
            assert np.array_equal(eng.posterior_block(i + (copies - 1) * nb), q0), i
