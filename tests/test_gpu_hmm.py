"""Parity of the CUDA recursions (through the C ABI) with the oracle and with the
reference-generated golden fixtures.  Tolerances are BASELINE.json's: log-likelihood
1e-9 relative (FP64), posterior 1e-7 absolute, Viterbi path bit-exact."""
import numpy as np
import pytest

import hmm_oracle as ho
import hmm_oracle_c as hoc
from conftest import golden, golden_models, golden_recursions

pytestmark = pytest.mark.gpu

LL_RTOL = 1e-9
POST_ATOL = 1e-7


def _tables(a, b, pi, V_lst):
    from itrails_b200.optimizer import viterbi_tables
    return viterbi_tables(a, b, pi, V_lst)


@pytest.mark.parametrize("fn", golden_recursions())
def test_golden_recursions(engine, fn):
    """CUDA vs the reference's own outputs (fixtures made by running the reference)."""
    g = golden(fn)
    m = golden(str(g["model_file"]))
    a, b, pi = m["a"], m["b"], m["pi"]
    n = int(g["n_blocks"])
    V_lst = [g[f"V_{i}"] for i in range(n)]
    engine.load_blocks(V_lst)
    engine.set_model(a, b, pi)
    tot, pb = engine.loglik(per_block=True)
    for i in range(n):
        ref = float(g[f"loglik_{i}"])
        assert abs(pb[0, i] - ref) <= LL_RTOL * abs(ref), (i, pb[0, i], ref)
    assert abs(tot[0] - float(g["loglik_total"])) <= LL_RTOL * abs(float(g["loglik_total"]))
    post = engine.split(engine.posterior())
    for i in range(n):
        assert np.abs(post[i] - g[f"post_{i}"]).max() <= POST_ATOL
    path = engine.split(engine.viterbi(*_tables(a, b, pi, V_lst)))
    for i in range(n):
        assert np.array_equal(path[i].astype(np.float64), g[f"vit_{i}"]), f"block {i}"


@pytest.mark.parametrize("fn", golden_models())
def test_every_state_count(engine, fn):
    """All golden models (K = 4, 5, 11, 13, 15, 21, 27): CUDA vs oracle on seeded data
    with N columns, ragged block lengths incl. 1, 31, 32, 33 and chunk boundaries."""
    m = golden(fn)
    a, b, pi = m["a"], m["b"], m["pi"]
    rng = np.random.default_rng(20261018)
    lens = [1, 2, 31, 32, 33, 64, 255, 256, 257, 513, 1000, 3001]
    V_lst = [ho.sample_block(a, b, pi, T, rng, p_n=0.03) for T in lens]
    V_lst[3][:] = 624          # all-N block (symbol NNNN)
    V_lst[5][::2] = 256        # AAAN every other column
    E = ho.emission_table(b)
    engine.load_blocks(V_lst)
    engine.set_model(a, b, pi)
    tot, pb = engine.loglik(per_block=True)
    ref = hoc.loglik_blocks(a, E, pi, V_lst)
    np.testing.assert_allclose(pb[0], ref, rtol=LL_RTOL, atol=1e-12)
    post = engine.split(engine.posterior())
    for p, r in zip(post, hoc.post_prob_blocks(a, E, pi, V_lst)):
        assert np.abs(p - r).max() <= POST_ATOL
        np.testing.assert_allclose(p.sum(1), 1.0, atol=1e-12)
    LA, LE, om0 = _tables(a, b, pi, V_lst)
    path = engine.split(engine.viterbi(LA, LE, om0))
    for p, r in zip(path, hoc.viterbi_blocks(LA, LE, om0, V_lst)):
        assert np.array_equal(p, r)


def _random_model(K, rng):
    a = rng.random((K, K)) ** 3 + 1e-6
    a += np.eye(K) * K
    a /= a.sum(1, keepdims=True)
    b = rng.random((K, 256)) ** 4 + 1e-9
    b /= b.sum(1, keepdims=True)
    pi = rng.random(K) + 0.1
    pi /= pi.sum()
    return a, b, pi


@pytest.mark.parametrize("K", [1, 2, 3, 7, 8, 9, 28, 31, 32, 33, 60, 64, 70, 97, 133, 200, 255])
def test_arbitrary_K(engine, K):
    """Register path (K <= 32, every padded width) and streamed path (K > 32)."""
    rng = np.random.default_rng(K)
    a, b, pi = _random_model(K, rng)
    lens = [1, 5, 40, 300, 777] if K > 64 else [1, 5, 40, 300, 777, 2500]
    V_lst = [rng.integers(0, 625, size=T) for T in lens]
    E = ho.emission_table(b)
    engine.load_blocks(V_lst)
    engine.set_model(a, b, pi)
    _, pb = engine.loglik(per_block=True)
    np.testing.assert_allclose(pb[0], hoc.loglik_blocks(a, E, pi, V_lst), rtol=LL_RTOL)
    post = engine.split(engine.posterior())
    for p, r in zip(post, hoc.post_prob_blocks(a, E, pi, V_lst)):
        assert np.abs(p - r).max() <= POST_ATOL
    LA, LE, om0 = _tables(a, b, pi, V_lst)
    path = engine.split(engine.viterbi(LA, LE, om0))
    for p, r in zip(path, hoc.viterbi_blocks(LA, LE, om0, V_lst)):
        assert np.array_equal(p, r)


def test_viterbi_exact_ties_take_first_index(engine):
    """Symmetric model (states 1 and 2 exchangeable) => exact ties; np.argmax and the
    kernel must both take the lowest index."""
    K = 3
    a = np.array([[0.9, 0.05, 0.05], [0.1, 0.8, 0.1], [0.1, 0.1, 0.8]])
    b = np.full((K, 256), 1.0 / 256)
    pi = np.array([0.2, 0.4, 0.4])
    V_lst = [np.zeros(600, dtype=np.int64), np.arange(300) % 625]
    engine.load_blocks(V_lst)
    engine.set_model(a, b, pi)
    LA, LE, om0 = _tables(a, b, pi, V_lst)
    path = engine.split(engine.viterbi(LA, LE, om0))
    for p, r in zip(path, ho.viterbi_wrapper(a, b, pi, V_lst)):
        assert np.array_equal(p.astype(np.float64), r)
    assert set(np.unique(path[0])) <= {0, 1}


def test_multi_set_loglik(engine):
    """Several parameter sets over the same blocks in one launch."""
    rng = np.random.default_rng(5)
    m = golden("model_2_2_example.npz")
    K = m["a"].shape[0]
    sets = [(m["a"], m["b"], m["pi"])] + [_random_model(K, rng) for _ in range(4)]
    V_lst = [ho.sample_block(m["a"], m["b"], m["pi"], T, rng) for T in (900, 100, 4000, 77)]
    engine.load_blocks(V_lst)
    engine.set_model(np.stack([s[0] for s in sets]), np.stack([s[1] for s in sets]),
                     np.stack([s[2] for s in sets]))
    tot, pb = engine.loglik(per_block=True)
    for k, (a, b, pi) in enumerate(sets):
        ref = hoc.loglik_blocks(a, ho.emission_table(b), pi, V_lst)
        np.testing.assert_allclose(pb[k], ref, rtol=LL_RTOL)
        assert abs(tot[k] - ref.sum()) <= LL_RTOL * abs(ref.sum())


def test_full_size_properties(engine):
    """BASELINE config-1/2 sized blocks (100 kb) — size-independent checks:
    (i) per-block log-likelihood equals the oracle's on a 100 kb block;
    (ii) block order / partition invariance: the total over blocks does not depend on
    how blocks are ordered; (iii) posterior rows sum to 1 and argmax-posterior agrees
    with Viterbi on the overwhelming majority of columns; (iv) Viterbi path re-scored on
    the host reproduces the kernel's final omega ordering (path is a valid argmax chain)."""
    m = golden("model_2_2_example.npz")
    a, b, pi = m["a"], m["b"], m["pi"]
    rng = np.random.default_rng(20261019)
    V_lst = [ho.sample_block(a, b, pi, T, rng) for T in (100_000, 65_537, 131_072)]
    E = ho.emission_table(b)
    engine.load_blocks(V_lst)
    engine.set_model(a, b, pi)
    tot, pb = engine.loglik(per_block=True)
    ref = hoc.loglik_blocks(a, E, pi, V_lst, n_threads=3)
    np.testing.assert_allclose(pb[0], ref, rtol=LL_RTOL)
    LA, LE, om0 = _tables(a, b, pi, V_lst)
    path = engine.split(engine.viterbi(LA, LE, om0))
    for p, r in zip(path, hoc.viterbi_blocks(LA, LE, om0, V_lst, n_threads=3)):
        assert np.array_equal(p, r)
    post = engine.split(engine.posterior())
    ref_post = hoc.post_prob_blocks(a, E, pi, V_lst[:1])
    assert np.abs(post[0] - ref_post[0]).max() <= POST_ATOL
    for p in post:
        np.testing.assert_allclose(p.sum(1), 1.0, atol=1e-12)
    # reversed block order gives the same per-block numbers bit for bit
    engine.load_blocks(V_lst[::-1])
    _, pb2 = engine.loglik(per_block=True)
    assert np.array_equal(pb2[0][::-1], pb[0])


def test_errors_are_loud(engine):
    with pytest.raises(ValueError):
        engine.load_blocks([np.array([0, 1, 700])])
    with pytest.raises(ValueError):
        engine.load_blocks([np.array([], dtype=np.int64)])
    with pytest.raises(ValueError):
        engine.set_model(np.eye(3), np.ones((3, 100)), np.ones(3))


def test_reference_style_wrappers(engine):
    """The drop-in functions a user of the reference calls."""
    import itrails_b200 as itb
    g = golden(golden_recursions()[0])
    m = golden(str(g["model_file"]))
    a, b, pi = m["a"], m["b"], m["pi"]
    V_lst = [g[f"V_{i}"] for i in range(int(g["n_blocks"]))]
    ll = itb.loglik_wrapper(a, b, pi, V_lst)
    assert type(ll) is float          # the reference returns a Python float (numba scalar)
    assert abs(ll - float(g["loglik_total"])) <= LL_RTOL * abs(float(g["loglik_total"]))
    assert itb.loglik_wrapper_par(a, b, pi, V_lst) == ll
    vit = itb.viterbi_wrapper(a, b, pi, V_lst)
    post = itb.post_prob_wrapper(a, b, pi, V_lst)
    for i in range(len(V_lst)):
        assert vit[i].dtype == np.float64 and np.array_equal(vit[i], g[f"vit_{i}"])
        assert post[i].shape == g[f"post_{i}"].shape
        assert np.abs(post[i] - g[f"post_{i}"]).max() <= POST_ATOL


@pytest.mark.parametrize("variant", ["stream", "stream8", "4warp", "1warp", "check", "check64"])
def test_viterbi_kernel_variants_agree(engine, variant, monkeypatch):
    """The speculate-and-verify sweep (few chains; 16- and 8-warp CTAs), the four-warps-per-chain sweep and the
    one-warp-per-chain sweep (many chains) are all bit-exact against the oracle, on
    near-tie-heavy data (t_A == t_B makes topologies 2 and 3 exchangeable)."""
    if variant == "1warp":
        monkeypatch.setenv("ITR_VITERBI_1WARP", "1")
    else:
        monkeypatch.setenv("ITR_VITERBI", variant)
    m = golden("model_3_3_example.npz")
    a, b, pi = m["a"], m["b"], m["pi"]
    rng = np.random.default_rng(99)
    V_lst = [ho.sample_block(a, b, pi, T, rng, p_n=0.02) for T in (20000, 1, 257, 5000, 31)]
    # uniform random symbols: the backpointers change at almost every column, so the
    # speculative sweep mispredicts (and repairs) constantly
    V_lst += [rng.integers(0, 625, size=T) for T in (3000, 2, 16, 17)]
    engine.load_blocks(V_lst)
    engine.set_model(a, b, pi)
    LA, LE, om0 = _tables(a, b, pi, V_lst)
    path = engine.split(engine.viterbi(LA, LE, om0))
    for p, r in zip(path, hoc.viterbi_blocks(LA, LE, om0, V_lst)):
        assert np.array_equal(p, r)


@pytest.mark.parametrize("variant", ["check", "stream", "stream8"])
@pytest.mark.parametrize("shift", [-7.0e10, -2.0 ** 36 + 3000.0, -1.0e300])
def test_viterbi_screen_refuses_the_hoist_shortcut_for_huge_omega(engine, variant, shift, monkeypatch):
    """The FP32 screen proves the emission hoist only for |omega| < 2^36.  With
    omega_0 shifted beyond that (or drifting across 2^36 inside the block, or near the top of the
    double range, where doubles are far apart and sums collapse into ties) the exponent test must
    send every column to the literal hoisting test; paths stay bit-exact against the C oracle."""
    monkeypatch.setenv("ITR_VITERBI", variant)
    m = golden("model_3_3_example.npz")
    a, b, pi = m["a"], m["b"], m["pi"]
    rng = np.random.default_rng(1234)
    V_lst = [ho.sample_block(a, b, pi, T, rng, p_n=0.02) for T in (6000, 700, 1, 33)]
    V_lst += [rng.integers(0, 625, size=900)]
    engine.load_blocks(V_lst)
    engine.set_model(a, b, pi)
    LA, LE, om0 = _tables(a, b, pi, V_lst)
    om0 = om0 + shift
    path = engine.split(engine.viterbi(LA, LE, om0))
    for p, r in zip(path, hoc.viterbi_blocks(LA, LE, om0, V_lst)):
        assert np.array_equal(p, r)


def test_many_blocks_and_async_overlap(engine):
    """700 short blocks (one-warp sweep, several chains per warp through the work queue)
    and the asynchronous mode: the three recursions enqueued back to back give the
    same results as the synchronous calls."""
    m = golden("model_2_2_example.npz")
    a, b, pi = m["a"], m["b"], m["pi"]
    rng = np.random.default_rng(3)
    V_lst = [ho.sample_block(a, b, pi, int(T), rng, p_n=0.02) for T in rng.integers(1, 400, size=700)]
    E = ho.emission_table(b)
    engine.load_blocks(V_lst)
    engine.set_model(a, b, pi)
    LA, LE, om0 = _tables(a, b, pi, V_lst)
    engine.set_async(True)
    try:
        tot = np.empty(1)
        ll = engine.loglik()
        path = engine.viterbi(LA, LE, om0)
        post = engine.posterior()
        engine.sync()
    finally:
        engine.set_async(False)
    ref = hoc.loglik_blocks(a, E, pi, V_lst)
    assert abs(ll[0] - ref.sum()) <= LL_RTOL * abs(ref.sum())
    for p, r in zip(engine.split(path), hoc.viterbi_blocks(LA, LE, om0, V_lst)):
        assert np.array_equal(p, r)
    for p, r in zip(engine.split(post), hoc.post_prob_blocks(a, E, pi, V_lst)):
        assert np.abs(p - r).max() <= POST_ATOL


def test_run_compressed_loglik_equals_plain_sweep(engine, monkeypatch):
    """The run-compressed forward sweep (invariant-column runs applied as precomputed
    powers) against the column-by-column sweep and the oracle: several parameter sets,
    N columns, short blocks; and data without a dominant class (plain sweep is chosen)."""
    m = golden("model_3_3_example.npz")
    a, b, pi = m["a"], m["b"], m["pi"]
    rng = np.random.default_rng(21)
    m2 = golden("model_3_3_asym1.npz")
    V_lst = [ho.sample_block(a, b, pi, T, rng, p_n=0.02) for T in (30000, 1, 2, 33, 64, 65, 4097, 12000)]
    V_lst[4][:] = 0                      # one block that is a single run
    engine.load_blocks(V_lst)
    A = np.stack([a, m2["a"]]); B = np.stack([b, m2["b"]]); PI = np.stack([pi, m2["pi"]])
    engine.set_model(A, B, PI)
    _, pb_runs = engine.loglik(per_block=True)
    monkeypatch.setenv("ITR_NO_RUNS", "1")
    _, pb_plain = engine.loglik(per_block=True)
    monkeypatch.delenv("ITR_NO_RUNS")
    np.testing.assert_allclose(pb_runs, pb_plain, rtol=1e-12)
    for k, (aa, bb, pp) in enumerate(((a, b, pi), (m2["a"], m2["b"], m2["pi"]))):
        ref = hoc.loglik_blocks(aa, ho.emission_table(bb), pp, V_lst)
        np.testing.assert_allclose(pb_runs[k], ref, rtol=LL_RTOL)
    # uniform random symbols: no class holds half of the columns
    V_rand = [rng.integers(0, 625, size=5000) for _ in range(3)]
    engine.load_blocks(V_rand)
    engine.set_model(a, b, pi)
    _, pb = engine.loglik(per_block=True)
    np.testing.assert_allclose(pb[0], hoc.loglik_blocks(a, ho.emission_table(b), pi, V_rand), rtol=LL_RTOL)


def test_every_run_length_and_table_slot(engine):
    """Runs of every length 1..70 of the dominant invariant column, at every alignment to
    the 32-column tiles, separated by single literal columns: each slot of the run-power
    table (2..8, 16, 24, 32 columns per step) and each two-step decomposition is used by the
    log-likelihood sweep and by both checkpoint sweeps of the posterior; against the oracle."""
    m = golden("model_3_3_example.npz")
    a, b, pi = m["a"], m["b"], m["pi"]
    rng = np.random.default_rng(33)
    literals = np.array([1, 17, 64 + 16 + 4 + 1, 255, 256, 300, 624, 27])      # non-invariant and N symbols
    blocks = []
    for shift in (0, 5, 13, 31):
        cols = [rng.integers(0, 256, size=shift + 1)]
        for n in rng.permutation(np.arange(1, 71)):
            cols.append(np.zeros(n, dtype=np.int64))                                # AAAA run
            cols.append(literals[rng.integers(0, len(literals), size=1)])
        blocks.append(np.concatenate(cols).astype(np.int64))
    blocks.append(np.zeros(39, dtype=np.int64))
    blocks.append(np.concatenate([np.zeros(24, dtype=np.int64), [300], np.zeros(7, dtype=np.int64)]))
    engine.load_blocks(blocks)
    engine.set_model(a, b, pi)
    _, pb = engine.loglik(per_block=True)
    np.testing.assert_allclose(pb[0], hoc.loglik_blocks(a, ho.emission_table(b), pi, blocks), rtol=LL_RTOL)
    post = engine.split(engine.posterior())
    for p_, r in zip(post, ho.post_prob_wrapper(a, b, pi, blocks)):
        assert np.abs(p_ - r).max() <= POST_ATOL


def test_out_of_range_symbol_is_rejected_by_the_library(engine):
    """list.index raises in the reference's maf_parser (read_data.py:113-115); through the
    C ABI a symbol > 624 is caught by the device-side range check of itr_load_blocks."""
    sym = np.zeros(5000, dtype=np.uint16)
    sym[4321] = 625
    off = np.array([0, 3000, 5000], dtype=np.int64)
    with pytest.raises(ValueError, match="column 4321"):
        engine.load_packed(sym, off)
    with pytest.raises(Exception):
        engine.loglik()                       # nothing is resident after a rejected load
    sym[4321] = 624
    engine.load_packed(sym, off)
    m = golden("model_2_2_example.npz")
    engine.set_model(m["a"], m["b"], m["pi"])
    ref = hoc.loglik_blocks(m["a"], ho.emission_table(m["b"]), m["pi"], [sym[:3000].astype(np.int64), sym[3000:].astype(np.int64)])
    np.testing.assert_allclose(engine.loglik(per_block=True)[1][0], ref, rtol=LL_RTOL)


def test_viterbi_stream_under_concurrency_is_stable(engine):
    """Regression: the decoupled Viterbi sweep synchronises its warps through shared-memory
    flags; a race in its roll-back phase only showed when other kernels (posterior,
    log-likelihood) ran beside it.  100 chains, the three recursions overlapped, repeated:
    every repetition must give the same bit-exact paths."""
    m = golden("model_3_3_example.npz")
    a, b, pi = m["a"], m["b"], m["pi"]
    rng = np.random.default_rng(2026)
    V_lst = [ho.sample_block(a, b, pi, int(T), rng, p_n=0.01) for T in rng.integers(15000, 30000, size=100)]
    engine.load_blocks(V_lst)
    engine.set_model(a, b, pi)
    LA, LE, om0 = _tables(a, b, pi, V_lst)
    ref = hoc.viterbi_blocks(LA, LE, om0, V_lst[:6])
    first = None
    engine.set_async(True)
    try:
        for rep in range(8):
            path = engine.viterbi(LA, LE, om0)
            engine.posterior(fetch=False)
            engine.loglik()
            engine.sync()
            if first is None:
                first = path.copy()
                for p, r in zip(engine.split(path)[:6], ref):
                    assert np.array_equal(p, r)
            else:
                assert np.array_equal(path, first), f"repetition {rep}"
    finally:
        engine.set_async(False)
