"""Pins at BASELINE.json's sizes that the small parity tests do not give:

* the GPU-built model must carry the 1e-9 log-likelihood bound at 10 Mb, not only on a few
  thousand columns: reference-run model (tests/golden/model_3_3_example.npz, produced by
  the reference's trans_emiss_calc) + CPU oracle sweep versus itr_build_model + itr_loglik
  on the config-2 alignment;
* config 5 at full size: 1 024 parameter sets x 10 Mb in one batched call, 32 of the sets
  spot-checked against the CPU oracle (NumPy model build + C forward sweep);
* config 2's optimiser loop on the 10 Mb alignment: a bounded Nelder-Mead run through the
  reference-style `optimizer` entry point, every evaluation in the history file checked
  for consistency and the first / best ones against the oracle;
* the streamed posterior (itr_posterior_stream) delivers exactly the rows of itr_posterior.
"""
import csv
import os

import numpy as np
import pytest
import yaml

import ctmc_oracle as co
import hmm_oracle as ho
import hmm_oracle_c as hoc
from conftest import ROOT, golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def config2(engine):
    import sys
    sys.path.insert(0, ROOT)
    import bench
    m = golden("model_3_3_example.npz")
    lengths = bench.workload_lengths("config2")
    V = bench.workload_blocks("config2", m["a"], m["b"], m["pi"], lengths, range(len(lengths)))
    assert sum(len(v) for v in V) == 10_000_000 and len(V) == 100
    return m, V


def test_gpu_built_model_keeps_the_loglik_bound_at_10mb(engine, config2):
    """north_star: log-likelihood within 1e-9 relative.  Reference-built (a, b, pi) through
    the CPU oracle on all 10 Mb versus the GPU-built model through the GPU sweep."""
    m, V = config2
    E = ho.emission_table(m["b"])
    want = hoc.loglik_blocks(m["a"], E, m["pi"], [v.astype(np.int64) for v in V], os.cpu_count() or 1)
    engine.load_blocks(V)
    engine.build_model(m["args"][None, :], 3, 3, fetch=False)
    tot, pb = engine.loglik(per_block=True)
    assert abs(tot[0] - want.sum()) <= 1e-9 * abs(want.sum())
    np.testing.assert_allclose(pb[0], want, rtol=1e-9)
    # and the reference-built model installed as is (isolates the sweep from the build)
    engine.set_model(m["a"], m["b"], m["pi"])
    tot2 = engine.loglik()
    assert abs(tot2[0] - want.sum()) <= 1e-9 * abs(want.sum())


def test_config5_1024_sets_spot_checked(engine, config2):
    """Config 5: 1 024 parameter sets x 10 Mb, batched model build + batched forward sweep.
    32 random sets are rebuilt with the NumPy oracle and swept with the C oracle over the
    first 12 blocks (1.2 Mb); the per-block GPU values of those sets must agree to 1e-9."""
    from test_gpu_model import _random_sets
    m, V = config2
    params = _random_sets(1024, 55)
    engine.load_blocks(V)
    engine.build_model(params, 3, 3, fetch=False)
    tot, pb = engine.loglik(per_block=True)
    assert tot.shape == (1024,) and pb.shape == (1024, 100) and np.isfinite(pb).all()
    np.testing.assert_allclose(tot, pb.sum(1), rtol=1e-12)
    sub = [v.astype(np.int64) for v in V[:12]]
    rng = np.random.default_rng(9)
    for s in rng.choice(1024, size=32, replace=False):
        a, b, pi = co.trans_emiss_calc(*params[s], 3, 3)[:3]
        want = hoc.loglik_blocks(a, ho.emission_table(b), pi, sub, os.cpu_count() or 1)
        np.testing.assert_allclose(pb[s, :12], want, rtol=1e-9, err_msg=f"set {s}")


def test_config2_nelder_mead_loop_on_10mb(engine, config2, tmp_path):
    """Config 2: the optimiser entry point on the 10 Mb alignment (bounded to 40 iterations
    here; the full loop is `tools/time_optimize.py`).  The history file must list every
    evaluation, the best-model file the best of them, and the first and the best
    evaluation are recomputed with the oracle."""
    from scipy.optimize import minimize
    from itrails_b200 import engine_cache
    from itrails_b200.optimizer import derive_times, model_args, optimization_wrapper
    from itrails_b200.workflows import prepare_optimize
    m, V = config2
    cfg = {"fixed_parameters": {"mu": 1e-8, "t_1": 240000, "t_2": 40000, "t_upper": 745069.3855, "N_ABC": 50000},
           "optimized_parameters": {"N_AB": [40000, 5000, 500000], "r": [2e-8, 1e-9, 1e-7]}}
    optim_variables, optim_list, bounds, fixed, case = prepare_optimize(cfg, 3, 3)
    res = str(tmp_path / "run")
    with open(res + ".best_model.yaml", "w") as fh:
        yaml.safe_dump({"fixed_parameters": {"mu": 1e-8}, "optimized_parameters": {},
                        "results": {"log_likelihood": None, "iteration": None}}, fh)
    engine_cache._ENGINE = engine
    try:
        out = minimize(optimization_wrapper, x0=optim_list,
                       args=(optim_variables, case, fixed.copy(), V, res, {"Nfeval": 0, "time": 0.0}),
                       method="Nelder-Mead", bounds=bounds, options={"maxiter": 40})
    finally:
        engine_cache._ENGINE = None
        engine_cache._LOADED = None
    hist = np.array([[float(x) for x in r] for r in csv.reader(open(res + ".optimization_history.csv"))])
    assert len(hist) == out.nfev and np.array_equal(hist[:, 0], np.arange(out.nfev))
    best = yaml.safe_load(open(res + ".best_model.yaml"))
    k = int(np.argmax(hist[:, 3]))
    assert best["results"]["iteration"] == k and abs(best["results"]["log_likelihood"] - hist[k, 3]) <= 1e-12 * abs(hist[k, 3])
    assert hist[k, 3] >= hist[0, 3] and abs(-out.fun - hist[k, 3]) <= 1e-12 * abs(out.fun)
    Vi = [v.astype(np.int64) for v in V]
    for row in (hist[0], hist[k]):
        d = fixed.copy()
        d.update(dict(zip(optim_variables, row[1:3])))
        derive_times(d, case)
        a, b, pi = co.trans_emiss_calc(*model_args(d), 3, 3)[:3]
        want = hoc.loglik_blocks(a, ho.emission_table(b), pi, Vi, os.cpu_count() or 1).sum()
        assert abs(row[3] - want) <= 1e-9 * abs(want)


def test_streamed_posterior_equals_posterior(engine):
    """itr_posterior_stream: every row arrives exactly once, in order, bit-identical to
    itr_posterior's result — with more pieces than ring slots, pieces that end at range
    boundaries, many blocks (ranged pass 2) and few blocks, K = 27 and K = 13, and a sink
    that stops the stream."""
    for fn, n_blocks in (("model_3_3_example.npz", 400), ("model_2_2_example.npz", 7)):
        m = golden(fn)
        a, b, pi = m["a"], m["b"], m["pi"]
        K = a.shape[0]
        rng = np.random.default_rng(5)
        V = [ho.sample_block(a, b, pi, int(T), rng, p_n=0.02) for T in rng.integers(1, 900, size=n_blocks)]
        engine.load_blocks(V)
        engine.set_model(a, b, pi)
        want = engine.posterior().copy()
        n = want.shape[0]
        slot_cols, n_slots = 1000, 3
        ring = np.zeros(n_slots * slot_cols * K)
        got = np.full_like(want, np.nan)
        seen = []

        def sink(col0, rows):
            seen.append((col0, len(rows)))
            got[col0:col0 + len(rows)] = rows
        engine.posterior_stream(ring, slot_cols, n_slots, sink)
        assert np.array_equal(got, want)
        assert [c for c, _ in seen] == sorted(c for c, _ in seen) and sum(k for _, k in seen) == n
        assert max(k for _, k in seen) <= slot_cols and len(seen) > n_slots
        # the result also stays on the device
        assert np.array_equal(engine.posterior_block(len(V) - 1), engine.split(want)[-1])

        class Stop(Exception):
            pass

        def bad_sink(col0, rows):
            raise Stop()
        with pytest.raises(Stop):
            engine.posterior_stream(ring, slot_cols, n_slots, bad_sink)
        engine.posterior_stream(ring, slot_cols, n_slots)            # no sink: ring only
        with pytest.raises(ValueError):
            engine.posterior_stream(ring[:10], slot_cols, n_slots)
