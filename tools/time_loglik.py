import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import itrails_b200 as itb
from itrails_b200 import synth
g = np.load(os.path.join(ROOT, "tests", "golden", "model_3_3_example.npz"))
rng = np.random.default_rng(1)
n, T = int(sys.argv[1]), int(sys.argv[2])
V = [synth.sample_block(g["a"], g["b"], g["pi"], T, rng) for _ in range(n)]
eng = itb.Engine(0); eng.load_blocks(V); eng.set_model(g["a"], g["b"], g["pi"])
for _ in range(3): ll = eng.loglik()
print(f"chains={n} T={T} loglik {eng.phase_ms('loglik'):.3f} ms = {eng.phase_ms('loglik')*1e-3*1.965e9/T:.0f} cycles/column/chain  ll={ll[0]:.4f}")
