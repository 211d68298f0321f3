"""Config 3 (100 Mb in 1 000 blocks at (5,5), K = 70) on one B200: log-likelihood and posterior
(kept in HBM) with the lock-step tensor-core sweeps (default for many blocks) and with the
one-CTA-per-chain sweeps (ITR_LOCKSTEP=0), device time from the library's phase events."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import itrails_b200 as itb
from itrails_b200 import synth

def tiled(V_base, copies):
    lens = np.array([len(v) for v in V_base] * copies, dtype=np.int64)
    off = np.zeros(len(lens) + 1, dtype=np.int64); off[1:] = np.cumsum(lens)
    return np.tile(np.concatenate(V_base).astype(np.uint16), copies), off

def timed(f, reps=3):
    f(); best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); f(); best = min(best, time.perf_counter() - t0)
    return best

n_int = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (5, 5)
copies = int(sys.argv[3]) if len(sys.argv) > 3 else 10
eng = itb.Engine(0)
rng = np.random.default_rng(4)
args = synth.example_model_args(n_int[1])
a, b, pi, _ = eng.build_model(args[None, :], *n_int)
a, b, pi = a[0], b[0], pi[0]
lens = synth.block_lengths(100, 10_000_000, rng)
V = synth.alignment(a, b, pi, lens, 40 + n_int[1])
sym, off = tiled(V, copies)
n, K = int(off[-1]), a.shape[0]
eng.load_packed(sym, off)
for mode in (None, "1", "0"):
    if mode is None: os.environ.pop("ITR_LOCKSTEP", None)
    else: os.environ["ITR_LOCKSTEP"] = mode
    for rot in ((None, "1") if mode == "1" else (None,)):
        if rot: os.environ["ITR_LOCKSTEP_NOROT"] = "1"
        else: os.environ.pop("ITR_LOCKSTEP_NOROT", None)
        l0 = eng.lockstep_launch_count
        t_ll = timed(lambda: eng.loglik())
        t_post = timed(lambda: eng.posterior(fetch=False))
        fl_ll, fl_post = n * (2.0 * K * K + 3 * K), n * (2.0 * (2 * K * K + 3 * K) + 3 * K)
        print(f"config3 {n_int} K={K} {n/1e6:.0f} Mb in {len(off)-1} blocks, ITR_LOCKSTEP={mode} norot={rot} "
              f"(lock-step launches {eng.lockstep_launch_count - l0}): loglik {t_ll*1e3:.1f} ms = {n/t_ll:.3g} col/s = "
              f"{fl_ll/t_ll/1e12:.2f} TFLOP/s; posterior {t_post*1e3:.1f} ms (device {eng.phase_ms('post_total'):.1f}) = "
              f"{n/t_post:.3g} col/s = {fl_post/t_post/1e12:.2f} TFLOP/s", flush=True)
