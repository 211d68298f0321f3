"""Per-step cost of the lock-step sweeps: n equal blocks of L columns at K = 70.
usage: time_lockstep.py L n_blocks [n_blocks ...]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import itrails_b200 as itb
from itrails_b200 import synth
L = int(sys.argv[1])
eng = itb.Engine(0)
a, b, pi, _ = eng.build_model(synth.example_model_args(5)[None, :], 5, 5)
V = synth.alignment(a[0], b[0], pi[0], np.full(8, L), 45)
os.environ["ITR_LOCKSTEP"] = "1"
def timed(f, reps=3):
    f(); best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); f(); best = min(best, time.perf_counter() - t0)
    return best
for nb in map(int, sys.argv[2:]):
    lens = np.full(nb, L, dtype=np.int64)
    off = np.zeros(nb + 1, dtype=np.int64); off[1:] = np.cumsum(lens)
    eng.load_packed(np.tile(np.concatenate(V).astype(np.uint16), nb // 8), off)
    for dbg in os.environ.get("DBG_LIST", "0").split(","):
        os.environ["ITR_LOCKSTEP_DBG"] = dbg
        t_ll = timed(lambda: eng.loglik())
        t_po = timed(lambda: eng.posterior(fetch=False))
        g0, g1 = (nb + 7) // 8, (nb + 3) // 4
        w0, w1 = max(1, -(-g0 // 296)), max(1, -(-g1 // 296))
        print(f"L={L} blocks={nb} dbg={dbg}: loglik {t_ll*1e3:.2f} ms ({g0} groups, {t_ll*1.965e9/L/w0:.0f} cycles/step/wave); "
              f"posterior {t_po*1e3:.2f} ms ({g1} groups, {t_po*1.965e9/L/w1:.0f} cycles/step/wave)", flush=True)
