"""The three console workflows end to end on a small synthetic MAF (GPU path): files
written, formats, and consistency with the wrapper functions."""
import csv
import os

import numpy as np
import pytest
import yaml

import hmm_oracle as ho
from conftest import golden

pytestmark = pytest.mark.gpu

SPECIES = ["hg38", "panTro5", "gorGor5", "ponAbe2"]


@pytest.fixture(scope="module")
def small_maf(tmp_path_factory):
    from itrails_b200 import synth
    m = golden("model_2_2_example.npz")
    rng = np.random.default_rng(8)
    V_lst = [ho.sample_block(m["a"], m["b"], m["pi"], T, rng, p_n=0.02) for T in (4000, 2500, 1, 3000)]
    d = tmp_path_factory.mktemp("wf")
    path = d / "small.maf"
    synth.write_maf(str(path), V_lst, SPECIES)
    return str(path), V_lst, str(d)


def test_optimize_then_decode(small_maf, engine):
    import itrails_b200 as itb
    from itrails_b200 import engine_cache, workflows
    maf, V_lst, d = small_maf
    cfg = {"fixed_parameters": {"mu": 1e-8, "t_1": 240000, "t_2": 40000, "t_upper": 745069.3855, "N_ABC": 50000},
           "optimized_parameters": {"N_AB": [40000, 5000, 500000], "r": [2e-8, 1e-9, 1e-7]},
           "settings": {"input_maf": None, "output_prefix": None, "n_cpu": 4, "method": "Nelder-Mead",
                        "species_list": SPECIES, "n_int_AB": 2, "n_int_ABC": 2}}
    cfg_path = os.path.join(d, "cfg.yaml")
    with open(cfg_path, "w") as fh:
        yaml.safe_dump(cfg, fh)
    prefix = os.path.join(d, "out", "run")
    engine_cache._ENGINE = engine            # share the session's GPU context
    res = workflows.optimize_main([cfg_path, "--input", maf, "--output", prefix])
    assert res.success or res.nit > 10
    hist = list(csv.reader(open(prefix + ".optimization_history.csv")))
    assert hist[0] == ["n_eval", "N_AB", "r", "loglik", "time"] and len(hist) > 20
    ll = np.array([float(r[3]) for r in hist[1:]])
    best = yaml.safe_load(open(prefix + ".best_model.yaml"))
    assert abs(best["results"]["log_likelihood"] - ll.max()) < 1e-9 * abs(ll.max())
    assert set(best["optimized_parameters"]) == {"N_AB", "r"}
    assert 5000 <= best["optimized_parameters"]["N_AB"] <= 500000
    start = yaml.safe_load(open(prefix + ".starting_params.yaml"))
    assert start["optimized_parameters"]["r"] == [2e-8, 1e-9, 1e-7]
    # the objective at the starting point equals the wrapper's log-likelihood
    from itrails_b200.workflows import prepare_decode
    dec = {"fixed_parameters": dict(cfg["fixed_parameters"], N_AB=40000, r=2e-8), "optimized_parameters": {},
           "settings": {"n_int_AB": 2, "n_int_ABC": 2}}
    fd, nab, nabc, _, _ = prepare_decode(dec)
    a, b, pi, hid, _ = itb.trans_emiss_calc(fd["t_A"], fd["t_B"], fd["t_C"], fd["t_2"], fd["t_upper"], fd["t_out"],
                                            fd["N_AB"], fd["N_ABC"], fd["r"], 2, 2)
    assert abs(itb.loglik_wrapper(a, b, pi, V_lst) - ll[0]) <= 1e-9 * abs(ll[0])

    # decoding from the best model, with reference coordinates
    out = workflows.viterbi_main(["--config-file", prefix + ".best_model.yaml", "--input", maf,
                                  "--output", prefix, "--reference", "hg38"])
    rows = list(csv.reader(open(out)))
    assert rows[0] == ["Block_idx", "position_start", "position_end", "most_likely_state"]
    assert {int(r[0]) for r in rows[1:]} == {0, 1, 2, 3}
    assert os.path.exists(prefix + ".hidden_states.csv")
    hs = list(csv.reader(open(prefix + ".hidden_states.csv")))
    assert len(hs) == 1 + 13 and hs[1][4] == "(0, 0, 0)"
    out = workflows.posterior_main(["--config-file", prefix + ".best_model.yaml", "--input", maf, "--output", prefix])
    assert os.path.exists(prefix + ".hidden_states_2.csv")
    rows = list(csv.reader(open(out)))
    assert rows[0][:3] == ["alignment_block_idx", "position_idx", "prob_state_0"] and len(rows[0]) == 2 + 13
    assert len(rows) == 1 + sum(len(v) for v in V_lst)
    probs = np.array(rows[1:4001], dtype=float)[:, 2:]
    np.testing.assert_allclose(probs.sum(1), 1.0, atol=1e-9)
    # the native streaming writer and the reference's csv.writer loop give the same bytes
    for extra in ([], ["--reference", "hg38"]):
        prefix2 = os.path.join(d, "out", "cmp" + str(len(extra)))
        out_native = workflows.posterior_main(["--config-file", prefix + ".best_model.yaml", "--input", maf,
                                               "--output", prefix2 + "_native"] + extra)
        os.environ["ITRAILS_PY_CSV"] = "1"
        try:
            out_py = workflows.posterior_main(["--config-file", prefix + ".best_model.yaml", "--input", maf,
                                               "--output", prefix2 + "_py"] + extra)
        finally:
            del os.environ["ITRAILS_PY_CSV"]
        assert open(out_native, "rb").read() == open(out_py, "rb").read()
    engine_cache._ENGINE = None
    engine_cache._LOADED = None


def test_loglik_sweep_equals_single_evaluations(small_maf, engine, tmp_path):
    """Batched objective (one batched model build + one multi-set forward sweep) against the
    per-point objective `optimization_wrapper` (optimizer.py:396-583)."""
    from itrails_b200 import engine_cache
    from itrails_b200.optimizer import loglik_sweep, optimization_wrapper
    from itrails_b200.workflows import prepare_optimize
    _maf, V_lst, _d = small_maf
    cfg = {"fixed_parameters": {"mu": 1e-8, "t_1": 240000, "t_2": 40000, "t_upper": 745069.3855, "N_ABC": 50000},
           "optimized_parameters": {"N_AB": [40000, 5000, 500000], "r": [2e-8, 1e-9, 1e-7]}}
    optim_variables, optim_list, bounds, fixed, case = prepare_optimize(cfg, 2, 2)
    engine_cache._ENGINE = engine
    try:
        rng = np.random.default_rng(4)
        lo = np.array([b[0] for b in bounds]); hi = np.array([b[1] for b in bounds])
        pts = np.vstack([optim_list] + [lo * (hi / lo) ** rng.random(len(lo)) for _ in range(6)])
        batched = loglik_sweep(pts, optim_variables, case, fixed, V_lst)
        assert batched.shape == (7,)
        res = str(tmp_path / "run")
        with open(res + ".best_model.yaml", "w") as fh:      # the workflow creates it before optimising
            yaml.safe_dump({"fixed_parameters": {"mu": 1e-8}, "optimized_parameters": {},
                            "results": {"log_likelihood": None, "iteration": None}}, fh)
        info = {"Nfeval": 0, "time": 0.0}
        single = np.array([-optimization_wrapper(p, optim_variables, case, fixed, V_lst, res, info) for p in pts])
        np.testing.assert_allclose(batched, single, rtol=1e-12)
        assert len(set(np.round(single, 3))) > 1
    finally:
        engine_cache._ENGINE = None
        engine_cache._LOADED = None


def test_batched_nelder_mead_workflow_matches_sequential(small_maf, engine):
    """`method: nelder-mead-batched` (speculative simplex, one batched objective call per
    iteration) against `method: Nelder-Mead` (scipy, one call per evaluation, as
    optimizer.py:623-637): same evaluations in the same order, same optimum."""
    import time
    from itrails_b200 import engine_cache, workflows
    maf, _V_lst, d = small_maf
    cfg = {"fixed_parameters": {"mu": 1e-8, "t_1": 240000, "t_2": 40000, "t_upper": 745069.3855, "N_ABC": 50000},
           "optimized_parameters": {"N_AB": [40000, 5000, 500000], "r": [2e-8, 1e-9, 1e-7]},
           "settings": {"input_maf": None, "output_prefix": None, "n_cpu": 4, "method": "Nelder-Mead",
                        "species_list": SPECIES, "n_int_AB": 2, "n_int_ABC": 2}}
    engine_cache._ENGINE = engine
    try:
        out, wall = {}, {}
        for method in ("Nelder-Mead", "nelder-mead-batched"):
            cfg["settings"]["method"] = method
            cfg_path = os.path.join(d, f"cfg_{method}.yaml")
            with open(cfg_path, "w") as fh:
                yaml.safe_dump(cfg, fh)
            prefix = os.path.join(d, "out_nm", method)
            t0 = time.perf_counter()
            res = workflows.optimize_main([cfg_path, "--input", maf, "--output", prefix])
            wall[method] = time.perf_counter() - t0
            hist = list(csv.reader(open(prefix + ".optimization_history.csv")))
            out[method] = (res, np.array([[float(v) for v in r[:4]] for r in hist[1:]]),
                           yaml.safe_load(open(prefix + ".best_model.yaml")))
        (rs, hs, bs), (rb, hb, bb) = out["Nelder-Mead"], out["nelder-mead-batched"]
        assert rb.nfev == rs.nfev and rb.nit == rs.nit and hb.shape == hs.shape
        np.testing.assert_array_equal(hb[:, 0], hs[:, 0])
        np.testing.assert_allclose(hb[:, 1:3], hs[:, 1:3], rtol=1e-12)        # parameter vectors
        np.testing.assert_allclose(hb[:, 3], hs[:, 3], rtol=1e-12)            # log-likelihoods
        np.testing.assert_allclose(rb.x, rs.x, rtol=1e-12)
        assert bb["results"]["iteration"] == bs["results"]["iteration"]
        assert rb.nbatch < rs.nfev
        print(f"nelder-mead {wall['Nelder-Mead']:.3f} s ({rs.nfev} evaluations), "
              f"batched {wall['nelder-mead-batched']:.3f} s ({rb.nbatch} calls, {rb.nspec} points)")
    finally:
        engine_cache._ENGINE = None
        engine_cache._LOADED = None
