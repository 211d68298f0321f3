"""Driver for ncu captures at config-4 shape (many blocks, K = 27): one recursion, run
`reps` times on a `scale` fraction of the 250 Mb alignment.

    ncu --set full --import-source on -k regex:posterior_tiles_mma -c 1 \
        python tools/prof_config4.py posterior 0.1
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import itrails_b200 as itb
from itrails_b200 import synth
from itrails_b200.optimizer import viterbi_tables

what = sys.argv[1] if len(sys.argv) > 1 else "posterior"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
eng = itb.Engine(0)
params = synth.example_model_args(3)[None, :]
a, b, pi, _ = eng.build_model(params, 3, 3)
a, b, pi = a[0], b[0], pi[0]
# all 2 500 blocks, each shortened to `scale` of its length: the many-chain launch shapes of config 4
lengths = np.maximum(64, (bench.workload_lengths("config4") * scale).astype(np.int64))
V = bench.workload_blocks("config4", a, b, pi, lengths, range(len(lengths)))
eng.load_blocks(V)
LA, LE, om0 = viterbi_tables(a, b, pi, V)
for _ in range(reps):
    if what == "posterior":
        eng.posterior(fetch=False)
        print("posterior ms", eng.phase_ms("post_total"), "tiles", eng.phase_ms("post_combine"),
              "fwd", eng.phase_ms("post_fwd"), "bwd", eng.phase_ms("post_bwd"))
    elif what == "viterbi":
        eng.viterbi(LA, LE, om0, fetch=False)
        print("viterbi fwd ms", eng.phase_ms("viterbi_fwd"), "trace", eng.phase_ms("viterbi_trace"))
    else:
        print("loglik", eng.loglik(), eng.phase_ms("loglik"))
