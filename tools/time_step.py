"""Times one resident step (loglik + Viterbi + posterior) sequentially and with the
recursions overlapped (async mode) on a config-2-like workload."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import itrails_b200 as itb
from itrails_b200.optimizer import viterbi_tables

n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 100
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
g = np.load(os.path.join(ROOT, "tests", "golden", "model_3_3_example.npz"))
a, b, pi = g["a"], g["b"], g["pi"]
rng = np.random.default_rng(1)
from itrails_b200 import synth
V = [synth.sample_block(a, b, pi, T, rng) for _ in range(n_chains)]
eng = itb.Engine(0)
eng.load_blocks(V)
eng.set_model(a, b, pi)
tabs = viterbi_tables(a, b, pi, V)
def step():
    ll = eng.loglik(); eng.viterbi(*tabs, fetch=False); eng.posterior(fetch=False); eng.sync(); return ll
for mode in (False, True, False, True):
    eng.set_async(mode)
    step(); step()
    t0 = time.perf_counter()
    for _ in range(5): ll = step()
    dt = (time.perf_counter() - t0) / 5
    print(f"async={mode}: {dt*1e3:.2f} ms/step  ll={ll[0]:.6f}  phases: " + " ".join(f"{k}={eng.phase_ms(k):.2f}" for k in ("loglik", "viterbi_fwd", "viterbi_trace", "post_fwd", "post_bwd", "post_combine", "post_total")))
