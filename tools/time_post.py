import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from itrails_b200 import _lib
if os.environ.get('ITR_LIB'):
    _lib.LIB_PATH = os.path.abspath(os.environ['ITR_LIB'])
import itrails_b200 as itb
from itrails_b200 import synth
g = np.load(os.path.join(ROOT, "tests", "golden", "model_3_3_example.npz"))
rng = np.random.default_rng(1)
n, T = int(sys.argv[1]), int(sys.argv[2])
V = [synth.sample_block(g["a"], g["b"], g["pi"], T, rng) for _ in range(n)]
eng = itb.Engine(0); eng.load_blocks(V); eng.set_model(g["a"], g["b"], g["pi"])
for _ in range(3): eng.posterior(fetch=False)
print(f"chains={n} T={T} posterior: sweeps {eng.phase_ms('post_fwd'):.2f} / {eng.phase_ms('post_bwd'):.2f} ms, tiles {eng.phase_ms('post_combine'):.3f} ms, total {eng.phase_ms('post_total'):.2f} ms")
