"""Wall-clock of the decode CLIs (itrails-posterior / itrails-viterbi) on a synthetic MAF:
MAF ingest (native reader), model build, recursion, result writer (native streaming CSV
writer vs the reference-style csv.writer loop on a sample)."""
import os, sys, time, tempfile, csv, io
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import itrails_b200 as itb
from itrails_b200 import synth, workflows, engine_cache
import yaml

total = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
SPECIES = ["hg38", "panTro5", "gorGor5", "ponAbe2"]
eng = engine_cache.get_engine()
args = synth.example_model_args(3)
a, b, pi, _ = eng.build_model(args[None, :], 3, 3)
rng = np.random.default_rng(11)
lens = synth.block_lengths(max(1, total // 100_000), total, rng)
V = synth.alignment(a[0], b[0], pi[0], lens, 12)
d = tempfile.mkdtemp(prefix="itr_cli_")
maf = os.path.join(d, "x.maf")
t0 = time.perf_counter(); synth.write_maf(maf, V, SPECIES); t1 = time.perf_counter()
print(f"synthetic MAF: {total} columns, {len(V)} blocks, {os.path.getsize(maf)/1e6:.0f} MB (written in {t1-t0:.1f} s)")
cfg = {"fixed_parameters": {"mu": 1e-8, "t_1": 240000, "t_2": 40000, "t_upper": 745069.3855, "N_ABC": 50000,
                            "N_AB": 50000, "r": 1e-8},
       "optimized_parameters": {}, "settings": {"species_list": SPECIES, "n_int_AB": 3, "n_int_ABC": 3}}
cfgp = os.path.join(d, "cfg.yaml"); yaml.safe_dump(cfg, open(cfgp, "w"))
t0 = time.perf_counter(); Vr = itb.maf_parser(maf, SPECIES); t1 = time.perf_counter()
print(f"maf_parser (native): {t1-t0:.2f} s = {total/(t1-t0):.3g} columns/s")
for what, main in (("viterbi", workflows.viterbi_main), ("posterior", workflows.posterior_main)):
    t0 = time.perf_counter()
    out = main(["--config-file", cfgp, "--input", maf, "--output", os.path.join(d, "run_" + what)])
    t1 = time.perf_counter()
    print(f"itrails-{what}: {t1-t0:.2f} s wall for {total} columns = {total/(t1-t0):.3g} columns/s; "
          f"output {os.path.getsize(out)/1e6:.0f} MB")
    os.remove(out)
# writer alone: native vs csv.writer loop (sample of 100k rows)
eng.load_blocks(V); eng.set_model(a[0], b[0], pi[0]); eng.posterior(fetch=False)
p = os.path.join(d, "w.csv")
t0 = time.perf_counter(); eng.write_posterior_csv(p); t1 = time.perf_counter()
print(f"native posterior writer: {t1-t0:.2f} s, {total/(t1-t0):.3g} rows/s, {os.path.getsize(p)/(t1-t0)/1e9:.2f} GB/s")
os.remove(p)
n = min(100_000, len(V[0])); blk = eng.posterior_block(0)[:n]
t0 = time.perf_counter()
with open(p, "w", newline="") as fh:
    w = csv.writer(fh)
    for i, row in enumerate(blk):
        w.writerow([0, i] + row.tolist())
t1 = time.perf_counter()
print(f"csv.writer loop (reference style): {n/(t1-t0):.3g} rows/s")
import shutil; shutil.rmtree(d)
