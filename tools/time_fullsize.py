"""BASELINE.json configs 3 and 4 at full size on one B200 (results stay in HBM):
config 4 = 2 500 blocks x ~100 kb = 250 Mb at K = 27 (Viterbi + posterior + loglik),
config 3 = 1 000 blocks = 100 Mb at (5,5), K = 70 (posterior + loglik)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import itrails_b200 as itb
from itrails_b200 import synth
from itrails_b200.optimizer import viterbi_tables

def tiled(V_base, copies):
    lens = np.array([len(v) for v in V_base] * copies, dtype=np.int64)
    off = np.zeros(len(lens) + 1, dtype=np.int64); off[1:] = np.cumsum(lens)
    return np.tile(np.concatenate(V_base).astype(np.uint16), copies), off

def timed(f, reps=2):
    f(); best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); f(); best = min(best, time.perf_counter() - t0)
    return best

eng = itb.Engine(0)
rng = np.random.default_rng(4)
for name, n_int, copies in (("config4", 3, 25), ("config3", 5, 10)):
    args = synth.example_model_args(n_int)
    a, b, pi, _ = eng.build_model(args[None, :], n_int, n_int)
    a, b, pi = a[0], b[0], pi[0]
    lens = synth.block_lengths(100, 10_000_000, rng)
    V = synth.alignment(a, b, pi, lens, 40 + n_int)
    sym, off = tiled(V, copies)
    n = int(off[-1])
    t0 = time.perf_counter(); eng.load_packed(sym, off); t_load = time.perf_counter() - t0
    K = a.shape[0]
    t_ll = timed(lambda: eng.loglik())
    t_post = timed(lambda: eng.posterior(fetch=False))
    line = f"{name}: {n/1e6:.0f} Mb in {len(off)-1} blocks, K={K}: upload {t_load*1e3:.0f} ms; loglik {t_ll*1e3:.1f} ms = {n/t_ll:.3g} col/s; posterior (kept in HBM, {n*K*8/1e9:.0f} GB) {t_post*1e3:.1f} ms = {n/t_post:.3g} col/s"
    if K <= 32:
        LA, LE, om0 = viterbi_tables(a, b, pi, V)
        om = np.tile(om0, (copies, 1))
        t_v = timed(lambda: eng.viterbi(LA, LE, om, fetch=False))
        line += f"; Viterbi incl. traceback {t_v*1e3:.1f} ms = {n/t_v:.3g} col/s"
    print(line, flush=True)
