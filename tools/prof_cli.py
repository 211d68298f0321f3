import os, sys, time, tempfile, cProfile, pstats
import numpy as np, yaml
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import itrails_b200 as itb
from itrails_b200 import synth, workflows, engine_cache
SPECIES = ["hg38", "panTro5", "gorGor5", "ponAbe2"]
eng = engine_cache.get_engine()
args = synth.example_model_args(3)
a, b, pi, _ = eng.build_model(args[None, :], 3, 3)
rng = np.random.default_rng(11)
V = synth.alignment(a[0], b[0], pi[0], synth.block_lengths(20, 2_000_000, rng), 12)
d = tempfile.mkdtemp(); maf = os.path.join(d, "x.maf"); synth.write_maf(maf, V, SPECIES)
cfg = {"fixed_parameters": {"mu": 1e-8, "t_1": 240000, "t_2": 40000, "t_upper": 745069.3855, "N_ABC": 50000, "N_AB": 50000, "r": 1e-8},
       "optimized_parameters": {}, "settings": {"species_list": SPECIES, "n_int_AB": 3, "n_int_ABC": 3}}
cfgp = os.path.join(d, "cfg.yaml"); yaml.safe_dump(cfg, open(cfgp, "w"))
pr = cProfile.Profile(); pr.enable()
workflows.viterbi_main(["--config-file", cfgp, "--input", maf, "--output", os.path.join(d, "v")])
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
