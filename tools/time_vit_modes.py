"""Viterbi forward sweep at the chain counts one GPU sees when config 4 (2 500 blocks of
50-150 kb) is split over 8, 4, 2, 1 GPUs, with every sweep kernel the library has:
device time of the sweep and of the traceback per mode.  usage: time_vit_modes.py [scale]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import itrails_b200 as itb
from itrails_b200 import synth, distributed
from itrails_b200.optimizer import viterbi_tables

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.2
eng = itb.Engine(0)
a, b, pi, _ = eng.build_model(synth.example_model_args(3)[None, :], 3, 3)
a, b, pi = a[0], b[0], pi[0]
lengths = np.maximum(64, (bench.workload_lengths("config4") * scale).astype(np.int64))
for world in (8, 4, 2, 1):
    ids = distributed.lpt_partition(lengths, world)[0]
    V = bench.workload_blocks("config4", a, b, pi, lengths, ids)
    eng.load_blocks(V)
    eng.set_model(a, b, pi)
    tabs = viterbi_tables(a, b, pi, V)
    ref = None
    for mode in (None, "stream16", "stream8", "4warp", "check", "check64"):
        if mode is None: os.environ.pop("ITR_VITERBI", None)
        else: os.environ["ITR_VITERBI"] = mode
        for _ in range(2): path = eng.viterbi(*tabs)
        if ref is None: ref = path
        ok = np.array_equal(path, ref)
        n = sum(len(v) for v in V)
        print(f"1/{world} of config 4 x {scale}: {len(V)} chains, {n/1e6:.1f} Mb, mode {mode}: sweep {eng.phase_ms('viterbi_fwd'):.2f} ms "
              f"({eng.phase_ms('viterbi_fwd')*1e-3*1.965e9/max(len(v) for v in V):.0f} cycles per column of the longest chain), "
              f"traceback {eng.phase_ms('viterbi_trace'):.2f} ms, same path as default: {ok}", flush=True)
