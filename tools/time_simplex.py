"""Config 2 (10 Mb, 100 blocks, n_int 3,3, six free parameters): wall time of N iterations
of the sequential simplex search (scipy Nelder-Mead, one objective call per evaluation, as
optimizer.py:623-637) against the speculative batched search (batched_simplex.py, one
batched call per iteration).  Usage: python tools/time_simplex.py [iterations]"""
import os, sys, time, tempfile, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yaml
from scipy.optimize import minimize
from itrails_b200 import synth, optimizer as opt
from itrails_b200.batched_simplex import minimize_neldermead_batched
from itrails_b200.workflows import prepare_optimize

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 60
g = np.load(os.path.join(ROOT, "tests", "golden", "model_3_3_example.npz"))
rng = np.random.default_rng(1)
V = synth.alignment(g["a"], g["b"], g["pi"], synth.block_lengths(100, 10_000_000, rng), 5)
cfg = {"fixed_parameters": {"mu": 1e-8},
       "optimized_parameters": {"N_AB": [40000, 5000, 500000], "N_ABC": [60000, 5000, 500000],
                                "t_1": [200000, 24000, 2400000], "t_2": [50000, 4000, 400000],
                                "t_upper": [600000, 74506.9385, 7450693.8556], "r": [2e-8, 1e-9, 1e-7]},
       "settings": {"n_int_AB": 3, "n_int_ABC": 3}}
names, start, bounds, fixed, case = prepare_optimize(cfg, 3, 3)
d = tempfile.mkdtemp()
best = {"fixed_parameters": {"mu": 1e-8}, "optimized_parameters": {},
        "results": {"log_likelihood": -float("inf"), "iteration": None}, "settings": {}}


def fresh(tag):
    res = os.path.join(d, tag)
    with open(res + ".best_model.yaml", "w") as fh:
        yaml.dump(best, fh)
    return res


opt.loglik_sweep(np.array([start] * 4), names, case, fixed, V)          # upload + warm both paths
opt.optimization_wrapper(np.array(start), names, case, fixed, V, fresh("warm"), {"Nfeval": 0, "time": time.time()})

with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    res = fresh("seq")
    t0 = time.perf_counter()
    r_seq = minimize(opt.optimization_wrapper, x0=start,
                     args=(names, case, fixed.copy(), V, res, {"Nfeval": 0, "time": time.time()}),
                     method="Nelder-Mead", bounds=bounds, options={"maxiter": iters})
    t_seq = time.perf_counter() - t0

    res = fresh("bat")
    hist = []
    t0 = time.perf_counter()
    info = {"Nfeval": 0, "time": time.time()}

    def consume(x, f):
        opt.write_list([info["Nfeval"]] + np.asarray(x).tolist() + [-f, time.time() - info["time"]],
                       res + ".optimization_history.csv")
        opt.update_best_model(res + ".best_model.yaml", names, x, -f, info["Nfeval"])
        info["Nfeval"] += 1
        hist.append(f)
    calls = []

    def batch(X):
        t = time.perf_counter()
        f = -opt.loglik_sweep(X, names, case, fixed, V)
        calls.append((time.perf_counter() - t) * 1e3)
        return f
    r_bat = minimize_neldermead_batched(batch, start,
                                        bounds=bounds, maxiter=iters, consume=consume)
    t_bat = time.perf_counter() - t0

same = r_seq.nfev == r_bat.nfev and np.allclose(r_seq.x, r_bat.x, rtol=1e-12, atol=0)
print(f"sequential Nelder-Mead: {r_seq.nit} iterations, {r_seq.nfev} evaluations, {t_seq:.3f} s "
      f"({t_seq / r_seq.nit * 1e3:.2f} ms/iteration)")
print(f"batched   Nelder-Mead: {r_bat.nit} iterations, {r_bat.nfev} evaluations consumed, "
      f"{r_bat.nbatch} batched calls ({r_bat.nspec} points), {t_bat:.3f} s "
      f"({t_bat / r_bat.nit * 1e3:.2f} ms/iteration)")
print(f"batched objective calls: median {np.median(calls):.2f} ms, max {max(calls):.2f} ms, sum {sum(calls):.0f} ms "
      f"(the rest of the wall time is the history / best-model files)")
print(f"same trajectory: {same}; -loglik {r_seq.fun:.6f} vs {r_bat.fun:.6f}; speed-up {t_seq / t_bat:.2f}x")
