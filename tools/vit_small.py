"""Small Viterbi run (for compute-sanitizer / ncu experiments on the sweep kernels)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import itrails_b200 as itb
from itrails_b200 import synth
from itrails_b200.optimizer import viterbi_tables
import hmm_oracle_c as hoc
g = np.load(os.path.join(ROOT, "tests", "golden", "model_3_3_example.npz"))
a, b, pi = g["a"], g["b"], g["pi"]
rng = np.random.default_rng(7)
n, T = int(sys.argv[1]), int(sys.argv[2])
V = [synth.sample_block(a, b, pi, T, rng) for _ in range(n)]
V += [rng.integers(0, 625, size=300)]
eng = itb.Engine(0); eng.load_blocks(V); eng.set_model(a, b, pi)
LA, LE, om0 = viterbi_tables(a, b, pi, V)
path = eng.split(eng.viterbi(LA, LE, om0))
ok = all(np.array_equal(p, r) for p, r in zip(path, hoc.viterbi_blocks(LA, LE, om0, V)))
print("viterbi small run: bit-exact" if ok else "viterbi small run: MISMATCH")
