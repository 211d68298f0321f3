import os, sys, time
import numpy as np
sys.path.insert(0, '/root/repo')
import itrails_b200 as itb
from itrails_b200 import synth
eng = itb.Engine(0)
args = synth.example_model_args(5)
for rep in range(3):
    t0=time.perf_counter(); a,b,pi,_=eng.build_model(args[None,:],5,5); t1=time.perf_counter()
    print(f"(5,5) build call {rep}: {1e3*(t1-t0):.2f} ms device {eng.phase_ms('model'):.2f}")
args = synth.example_model_args(7)
for rep in range(2):
    t0=time.perf_counter(); a,b,pi,_=eng.build_model(args[None,:],7,7); t1=time.perf_counter()
    print(f"(7,7) K={a.shape[1]} build call {rep}: {1e3*(t1-t0):.2f} ms device {eng.phase_ms('model'):.2f}")
