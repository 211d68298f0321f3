"""One lock-step log-likelihood and one lock-step posterior on config 3 (ncu target)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import itrails_b200 as itb
from itrails_b200 import synth
copies = int(sys.argv[1]) if len(sys.argv) > 1 else 10
eng = itb.Engine(0)
rng = np.random.default_rng(4)
a, b, pi, _ = eng.build_model(synth.example_model_args(5)[None, :], 5, 5)
lens = synth.block_lengths(100, 10_000_000, rng)
V = synth.alignment(a[0], b[0], pi[0], lens, 45)
lens = np.array([len(v) for v in V] * copies, dtype=np.int64)
off = np.zeros(len(lens) + 1, dtype=np.int64); off[1:] = np.cumsum(lens)
eng.load_packed(np.tile(np.concatenate(V).astype(np.uint16), copies), off)
os.environ["ITR_LOCKSTEP"] = "1"
print(eng.loglik()[0])
eng.posterior(fetch=False)
print(eng.phase_ms("post_total"))
