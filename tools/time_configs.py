"""Timings of the other BASELINE.json configurations (scaled to fit a short run):
config 3 (finer discretisation, K = 70, posterior), config 5 (batched parameter sweep)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import itrails_b200 as itb
from itrails_b200 import synth
eng = itb.Engine(0)
args = synth.example_model_args(3)
rng = np.random.default_rng(3)

# ---- config 5 (scaled): n_sets parameter sets x 10 Mb
a, b, pi, _ = eng.build_model(args[None, :], 3, 3)
lens = synth.block_lengths(100, 10_000_000, rng)
V = synth.alignment(a[0], b[0], pi[0], lens, 7)
eng.load_blocks(V)
for n_sets in (1, 64, 256, 1024):
    P = np.repeat(args[None, :], n_sets, axis=0) * np.exp(rng.uniform(-0.2, 0.2, size=(n_sets, 9)))
    P[:, 2] = (P[:, 0] + P[:, 1]) / 2 + P[:, 3]
    eng.build_model(P, 3, 3, fetch=False); eng.loglik()
    t0 = time.perf_counter(); eng.build_model(P, 3, 3, fetch=False); t1 = time.perf_counter()
    ll = eng.loglik(); t2 = time.perf_counter()
    print(f"config5-like: {n_sets} sets x 10 Mb: build {1e3*(t1-t0):.2f} ms (device {eng.phase_ms('model'):.2f}), loglik {1e3*(t2-t1):.1f} ms "
          f"(device {eng.phase_ms('loglik'):.1f}) => {n_sets*1e7/(t2-t1):.3g} column-evaluations/s, {n_sets/(t2-t0):.1f} objective evaluations/s")

# ---- config 3 (scaled): K = 70 posterior, 10 Mb
t0 = time.perf_counter(); a, b, pi, _ = eng.build_model(args[None, :], 5, 5); t1 = time.perf_counter()
print(f"config3-like: (5,5) K={a.shape[1]} model build {1e3*(t1-t0):.2f} ms (device {eng.phase_ms('model'):.2f})")
V = synth.alignment(a[0], b[0], pi[0], lens, 9)
eng.load_blocks(V)
eng.posterior(fetch=False)
t0 = time.perf_counter(); eng.posterior(fetch=False); t1 = time.perf_counter()
ll = eng.loglik()
print(f"config3-like: posterior of 10 Mb at K=70: {1e3*(t1-t0):.1f} ms => {1e7/(t1-t0):.3g} columns/s; loglik {eng.phase_ms('loglik'):.1f} ms")
