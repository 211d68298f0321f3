"""Quick per-kernel timing on a config-2-like workload (one chain per SM), used to
compare kernel variants: prints ns and SM cycles per column per chain."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from itrails_b200 import _lib
if os.environ.get("ITR_LIB"):
    _lib.LIB_PATH = os.environ["ITR_LIB"]
import itrails_b200 as itb
from itrails_b200.optimizer import viterbi_tables

n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 100
T = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
g = np.load(os.path.join(ROOT, "gpurun_tmp", "model_3_3_oracle.npz"))
a, b, pi = g["a"], g["b"], g["pi"]
rng = np.random.default_rng(1)
V = [rng.integers(0, 256, size=T) for _ in range(n_chains)]
eng = itb.Engine(0)
eng.load_blocks(V)
eng.set_model(a, b, pi)
tabs = viterbi_tables(a, b, pi, V)
for _ in range(2):
    eng.loglik(); eng.viterbi(*tabs, fetch=False); eng.posterior(fetch=False)
ghz = 1.965
res = {}
for name, ph in (("loglik", "loglik"), ("viterbi_fwd", "viterbi_fwd"), ("post_fwd", "post_fwd"), ("post_bwd", "post_bwd"), ("post_total", "post_total")):
    res[name] = eng.phase_ms(ph)
print(os.environ.get("ITR_LIB", "default"), f"chains={n_chains} T={T}", " ".join(f"{k}={v:.2f}ms({v*1e6/T*ghz:.0f}cyc/col)" for k, v in res.items()))
