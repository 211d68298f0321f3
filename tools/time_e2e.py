"""Posterior with a pinned host destination: one group vs grouped early download."""
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
for rep in range(3):
    t0 = time.perf_counter(); eng.posterior(out=post); t1 = time.perf_counter()
    print(f"posterior->pinned host: {1e3*(t1-t0):.1f} ms  (post_total {eng.phase_ms('post_total'):.1f} ms, fwd {eng.phase_ms('post_fwd'):.1f})  groups={'1' if os.environ.get('ITR_POST_ONE_GROUP') else '8'}")
t0 = time.perf_counter(); eng.posterior(fetch=False); t1 = time.perf_counter()
print(f"posterior resident: {1e3*(t1-t0):.1f} ms")
from itrails_b200.optimizer import viterbi_tables
tabs = viterbi_tables(a, b, pi, V)
path = torch.empty(n, dtype=torch.uint8, pin_memory=True).numpy()
def e2e(order):
    eng.set_async(True)
    for what in order:
        if what == "p": eng.posterior(out=post)
        if what == "v": eng.viterbi(*tabs, out=path)
        if what == "l": ll = eng.loglik()
    eng.sync(); eng.set_async(False)
for order in ("pvl", "vlp", "vp", "pv", "p", "v", "pl"):
    e2e(order)
    t0 = time.perf_counter(); e2e(order); e2e(order); t1 = time.perf_counter()
    print(f"order {order}: {1e3*(t1-t0)/2:.1f} ms   vit_fwd {eng.phase_ms('viterbi_fwd'):.1f} post_fwd {eng.phase_ms('post_fwd'):.1f} post_total {eng.phase_ms('post_total'):.1f} loglik {eng.phase_ms('loglik'):.1f}")
sym, off = itb.Engine.pack_blocks(V)
sym_pin = torch.empty(len(sym), dtype=torch.uint16, pin_memory=True).numpy(); sym_pin[:] = sym
for rep in range(4):
    ts = [time.perf_counter()]
    eng.load_packed(sym_pin, off); ts.append(time.perf_counter())
    eng.set_model(a, b, pi); ts.append(time.perf_counter())
    eng.set_async(True)
    eng.posterior(out=post); ts.append(time.perf_counter())
    eng.viterbi(*tabs, out=path); ts.append(time.perf_counter())
    ll = eng.loglik(); ts.append(time.perf_counter())
    eng.sync(); ts.append(time.perf_counter()); eng.set_async(False)
    d = np.diff(ts) * 1e3
    print("bench-style e2e: total %.1f ms = load %.1f set_model %.1f enq_post %.1f enq_vit %.1f enq_ll %.1f sync %.1f" % (d.sum(), *d))
