"""The bench's resident step (model build + log-likelihood || Viterbi || posterior) on rank 0's
share of config 4 when it is LPT-split over 8, 4, 2, 1 GPUs — on ONE GPU, so that the strong-
scaling behaviour of a rank can be studied without a multi-GPU box.
usage: time_share_step.py [scale] [modes: comma list of ITR_VITERBI values, '-' = default]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import itrails_b200 as itb
from itrails_b200 import synth, distributed

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["-"]
worlds = [int(w) for w in sys.argv[3].split(",")] if len(sys.argv) > 3 else [8, 4, 2, 1]
eng = itb.Engine(0)
params = synth.example_model_args(3)[None, :]
a, b, pi, _ = eng.build_model(params, 3, 3)
a, b, pi = a[0], b[0], pi[0]
lengths = np.maximum(64, (bench.workload_lengths("config4") * scale).astype(np.int64))
for world in worlds:
    ids = distributed.lpt_partition(lengths, world)[0]
    V = bench.workload_blocks("config4", a, b, pi, lengths, ids)
    run = bench.Runner(eng, params, 3, 3, a, b, pi, V, 1, 0)
    run.load()
    for mode in modes:
        if mode == "-": os.environ.pop("ITR_VITERBI", None)
        else: os.environ["ITR_VITERBI"] = mode
        for _ in range(3): run.step_resident()
        ts = []
        for _ in range(5):
            t0 = time.perf_counter(); run.step_resident(); ts.append(time.perf_counter() - t0)
        ph = {p: round(eng.phase_ms(p), 2) for p in bench.PHASES}
        print(f"1/{world} of config 4 x {scale}: {run.nblk} chains, {run.ncol/1e6:.1f} Mb, ITR_VITERBI={mode}: step {np.median(ts)*1e3:.2f} ms; {ph}", flush=True)
