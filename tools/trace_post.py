"""Per-block timeline of the posterior download (ITR_POST_TRACE=1)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import itrails_b200 as itb
from itrails_b200 import synth
g = np.load(os.path.join(ROOT, "tests", "golden", "model_3_3_example.npz"))
a, b, pi = g["a"], g["b"], g["pi"]
rng = np.random.default_rng(1)
lens = synth.block_lengths(100, 10_000_000, rng)
V = synth.alignment(a, b, pi, lens, 5)
eng = itb.Engine(0)
eng.load_blocks(V); eng.set_model(a, b, pi)
n = eng.n_columns
post = torch.empty(n * 27, dtype=torch.float64, pin_memory=True).numpy().reshape(n, 27)
eng.posterior(out=post); eng.posterior(out=post)
os.environ["ITR_POST_TRACE"] = "1"
t0 = time.perf_counter(); eng.posterior(out=post); t1 = time.perf_counter()
print(f"traced call {1e3*(t1-t0):.1f} ms")
from itrails_b200.optimizer import viterbi_tables
tabs = viterbi_tables(a, b, pi, V)
path = torch.empty(n, dtype=torch.uint8, pin_memory=True).numpy()
del os.environ["ITR_POST_TRACE"]
def pv():
    eng.set_async(True); eng.posterior(out=post)
    if "v" in sys.argv[1]: eng.viterbi(*tabs, out=path)
    if "l" in sys.argv[1]: eng.loglik()
    eng.sync(); eng.set_async(False)
pv(); pv()
os.environ["ITR_POST_TRACE"] = "1"
print("---- pv", file=sys.stderr, flush=True)
t0 = time.perf_counter(); pv(); t1 = time.perf_counter()
print(f"traced {sys.argv[1]} {1e3*(t1-t0):.1f} ms vit_fwd {eng.phase_ms('viterbi_fwd'):.1f} loglik {eng.phase_ms('loglik'):.1f}")
