import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from itrails_b200 import _lib
if os.environ.get('ITR_LIB'):
    _lib.LIB_PATH = os.path.abspath(os.environ['ITR_LIB'])
import itrails_b200 as itb
from itrails_b200 import synth
from itrails_b200.optimizer import viterbi_tables
g = np.load(os.path.join(ROOT, "tests", "golden", "model_3_3_example.npz"))
rng = np.random.default_rng(1)
n, T = int(sys.argv[1]), int(sys.argv[2])
V = [synth.sample_block(g["a"], g["b"], g["pi"], T, rng) for _ in range(n)]
eng = itb.Engine(0); eng.load_blocks(V); eng.set_model(g["a"], g["b"], g["pi"])
tabs = viterbi_tables(g["a"], g["b"], g["pi"], V)
for _ in range(3): eng.viterbi(*tabs, fetch=False)
print(f"chains={n} T={T} viterbi_fwd {eng.phase_ms('viterbi_fwd'):.3f} ms = {eng.phase_ms('viterbi_fwd')*1e-3*1.965e9/T:.0f} cycles/column/chain")
