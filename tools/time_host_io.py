"""Host-side rows of the path on this machine's CPU cores (no GPU needed): the native MAF reader
(N1) and the native posterior CSV writer on a host matrix (N2), next to the reference's own
csv.writer loop (workflow_posterior.py:697-716) on a sample."""
import csv, os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from itrails_b200 import read_data, synth, _lib as L

mb = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
rng = np.random.default_rng(1)
g = np.load(os.path.join(ROOT, "tests", "golden", "model_3_3_example.npz"))
a, b, pi = g["a"], g["b"], g["pi"]
n_blocks = max(1, int(mb * 1e6 / 100000))
V = [synth.sample_block(a, b, pi, 100000, rng) for _ in range(n_blocks)]
sp = ["hg38", "panTro5", "gorGor5", "ponAbe2"]
with tempfile.TemporaryDirectory() as d:
    maf = os.path.join(d, "x.maf")
    synth.write_maf(maf, V, species=tuple(sp))
    size = os.path.getsize(maf)
    t0 = time.perf_counter(); sym, off, coord, coff = read_data.read_maf(maf, sp, ref="hg38"); t1 = time.perf_counter()
    ncol = int(off[-1])
    print(f"MAF reader: {ncol/1e6:.1f} Mb in {n_blocks} blocks, {size/1e6:.0f} MB file, {t1-t0:.3f} s = {ncol/(t1-t0):.3e} columns/s, "
          f"{size/(t1-t0)/1e9:.2f} GB/s ({os.cpu_count()} cores), symbols + coordinates")
    K = 27
    rows = 2_000_000
    post = rng.random((rows, K)); post /= post.sum(1, keepdims=True)
    lib = L.load()
    import ctypes
    out = os.path.join(d, "p.csv")
    nb = 20
    offs = np.linspace(0, rows, nb + 1).astype(np.int64); pos = np.arange(rows, dtype=np.int64)
    t0 = time.perf_counter()
    rc = lib.itr_csv_posterior_host(out.encode(), K, nb, L.as_ptr(offs, ctypes.c_int64), L.as_ptr(pos, ctypes.c_int64),
                                    L.as_ptr(post, ctypes.c_double), os.cpu_count())
    t1 = time.perf_counter()
    assert rc == 0, rc
    sz = os.path.getsize(out)
    print(f"native CSV writer: {rows} rows x {K} in {t1-t0:.3f} s = {rows/(t1-t0):.3e} rows/s, {sz/(t1-t0)/1e9:.2f} GB/s of text")
    n_py = 100_000
    t0 = time.perf_counter()
    with open(os.path.join(d, "q.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["Block_idx", "position"] + [str(i) for i in range(K)])
        for i in range(n_py):
            w.writerow([0, i] + post[i].tolist())
    t1 = time.perf_counter()
    print(f"reference's csv.writer loop: {n_py} rows in {t1-t0:.3f} s = {n_py/(t1-t0):.3e} rows/s")
