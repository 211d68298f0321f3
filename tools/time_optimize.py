"""Objective evaluations per second of the itrails-optimize loop on config 2 (10 Mb, 100
blocks, n_int 3,3): the reference's optimization_wrapper signature, GPU path."""
import os, sys, time, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yaml
import itrails_b200 as itb
from itrails_b200 import synth, optimizer as opt
from itrails_b200.workflows import prepare_optimize

g = np.load(os.path.join(ROOT, "tests", "golden", "model_3_3_example.npz"))
rng = np.random.default_rng(1)
lens = synth.block_lengths(100, 10_000_000, rng)
V = synth.alignment(g["a"], g["b"], g["pi"], lens, 5)
cfg = {"fixed_parameters": {"mu": 1e-8},
       "optimized_parameters": {"N_AB": [50000, 5000, 500000], "N_ABC": [50000, 5000, 500000],
                                "t_1": [240000, 24000, 2400000], "t_2": [40000, 4000, 400000],
                                "t_upper": [745069.3855, 74506.9385, 7450693.8556], "r": [1e-8, 1e-9, 1e-7]},
       "settings": {"n_int_AB": 3, "n_int_ABC": 3}}
names, start, bounds, fixed, case = prepare_optimize(cfg, 3, 3)
d = tempfile.mkdtemp()
res = os.path.join(d, "run")
with open(res + ".best_model.yaml", "w") as fh:
    yaml.dump({"fixed_parameters": {"mu": 1e-8}, "optimized_parameters": {}, "results": {"log_likelihood": -float("inf"), "iteration": None}, "settings": {}}, fh)
info = {"Nfeval": 0, "time": time.time()}
x = np.array(start)
opt.optimization_wrapper(x, names, case, fixed, V, res, info)     # uploads the alignment
t0 = time.perf_counter()
n = 50
for i in range(n):
    opt.optimization_wrapper(x * (1 + 0.01 * np.sin(i + np.arange(len(x)))), names, case, fixed, V, res, info)
dt = (time.perf_counter() - t0) / n
print(f"objective evaluation (model build + 10 Mb log-likelihood + history/best-model files): {dt*1e3:.2f} ms  => {1/dt:.1f} evaluations/s")
