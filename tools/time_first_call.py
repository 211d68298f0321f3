"""Cold-start cost of each entry point in a fresh process (first vs second call)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
t0 = time.perf_counter(); import itrails_b200 as itb; from itrails_b200 import synth
from itrails_b200.optimizer import viterbi_tables
print(f"import {time.perf_counter()-t0:.2f} s")
g = np.load(os.path.join(ROOT, "tests", "golden", "model_3_3_example.npz"))
a, b, pi = g["a"], g["b"], g["pi"]
rng = np.random.default_rng(1)
V = [synth.sample_block(a, b, pi, 100000, rng) for _ in range(20)]
def T(label, f):
    t0 = time.perf_counter(); r = f(); print(f"{label}: {1e3*(time.perf_counter()-t0):.1f} ms", flush=True); return r
eng = T("Engine()", lambda: itb.Engine(0))
T("load_blocks", lambda: eng.load_blocks(V)); T("load_blocks again", lambda: eng.load_blocks(V))
T("set_model", lambda: eng.set_model(a, b, pi)); T("set_model again", lambda: eng.set_model(a, b, pi))
T("loglik", lambda: eng.loglik()); T("loglik again", lambda: eng.loglik())
tabs = T("viterbi_tables", lambda: viterbi_tables(a, b, pi, V))
T("viterbi", lambda: eng.viterbi(*tabs)); T("viterbi again", lambda: eng.viterbi(*tabs))
T("posterior(fetch=False)", lambda: eng.posterior(fetch=False)); T("posterior again", lambda: eng.posterior(fetch=False))
args = synth.example_model_args(3)
T("build_model", lambda: eng.build_model(args[None, :], 3, 3)); T("build_model again", lambda: eng.build_model(args[None, :], 3, 3))
