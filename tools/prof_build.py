"""Batched model build (config 5's builder half) for ncu: `n_sets` parameter sets at (3,3).
    ncu --metrics gpu__time_duration.sum --clock-control none python tools/prof_build.py 1024
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import itrails_b200 as itb
from itrails_b200 import synth
n_sets = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
eng = itb.Engine(0)
rng = np.random.default_rng(3)
base = synth.example_model_args(3)[None, :]
params = base * (1.0 + 0.02 * rng.standard_normal((n_sets, 9)))
for _ in range(reps):
    eng.build_model(params, 3, 3, fetch=False)
    print("build ms", eng.phase_ms("model"))
