"""Which recursion slows which inside the bench's resident step: the step with subsets of
{log-likelihood, Viterbi, posterior} enqueued, per-phase device times of each run.
usage: time_overlap.py [world: rank 0's share of config 4 split over `world` GPUs]"""
import os, sys, time, itertools
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import itrails_b200 as itb
from itrails_b200 import synth, distributed

world = int(sys.argv[1]) if len(sys.argv) > 1 else 1
eng = itb.Engine(0)
params = synth.example_model_args(3)[None, :]
a, b, pi, _ = eng.build_model(params, 3, 3)
a, b, pi = a[0], b[0], pi[0]
lengths = bench.workload_lengths("config4")
ids = distributed.lpt_partition(lengths, world)[0]
V = bench.workload_blocks("config4", a, b, pi, lengths, ids)
run = bench.Runner(eng, params, 3, 3, a, b, pi, V, 1, 0)
run.load()
for combo in ["p", "v", "l", "pv", "pl", "vl", "pvl"]:
    def step():
        eng.set_async(True)
        eng.build_model(run.params, 3, 3, fetch=False)
        if "v" in combo: eng.viterbi(run.log_a, run.log_E, run.omega0, fetch=False)
        if "p" in combo: eng.posterior(fetch=False)
        if "l" in combo: eng.loglik()
        eng.sync()
        eng.set_async(False)
    for _ in range(2): step()
    ts = []
    for _ in range(4):
        t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
    ph = {p: round(eng.phase_ms(p), 2) for p in bench.PHASES}
    print(f"{combo:>4s}: step {np.median(ts)*1e3:7.2f} ms  {ph}", flush=True)
