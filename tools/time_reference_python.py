"""Times the UNMODIFIED reference (trails-phylogeny/itrails, pure Python + numba) on this
host's cores: the recursions of optimizer.py:146-377 through the reference's own
functions, imported from baseline/_ref (installed with
`pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>`;
git-ignored, travels to the GPU box).  Biopython is absent from the image: a stub package
(oracle/_stubs/Bio) satisfies `import Bio.AlignIO`; no MAF is parsed here.  The model
(a, b, pi) comes from a fixture produced by the reference's trans_emiss_calc (420-670 s
per call at (3,3), SURVEY 6 — not repeated here).

    python tools/time_reference_python.py [columns of the block = 100000] > profiles/reference_python_r2.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
if not os.path.isdir(os.path.join(REF, "itrails")):
    print(json.dumps({"unavailable": "baseline/_ref/itrails is not installed"}))
    sys.exit(0)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_stubs"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(REF, ".nbcache"))

T = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
t_imp = time.perf_counter()
import itrails.ncpu as ncpu                      # noqa: E402
import itrails.optimizer as ro                   # noqa: E402
from itrails.read_data import get_idx_state      # noqa: E402
t_imp = time.perf_counter() - t_imp
cores = os.cpu_count() or 1
ncpu.update_n_cpu(cores)

import bench                                     # noqa: E402
g = np.load(os.path.join(ROOT, "tests", "golden", "model_3_3_example.npz"))
a, b, pi = g["a"], g["b"], g["pi"]
V = bench.workload_blocks("config1", a, b, pi, np.array([T]), [0])[0].astype(np.int64)
out = {"what": "the reference's own Python/numba functions on this host (config 1: one block, K = 27)",
       "columns": T, "cores": cores, "import_s": t_imp}

t0 = time.perf_counter()
order = [get_idx_state(i) for i in range(625)]            # optimizer.py:54 — rebuilt on every wrapper call
out["order_rebuild_s"] = time.perf_counter() - t0
import numba                                      # noqa: E402
order_nb = numba.typed.List(order)
ro.forward_loglik(a, b, pi, V[:1000], order_nb)    # JIT
t0 = time.perf_counter()
ll = ro.forward_loglik(a, b, pi, V, order_nb)
dt = time.perf_counter() - t0
out["forward_loglik"] = {"columns_per_s": T / dt, "s": dt, "loglik": float(ll)}
ro.post_prob(a, b, pi, V[:1000], order_nb)
t0 = time.perf_counter()
post = ro.post_prob(a, b, pi, V, order_nb)
dt = time.perf_counter() - t0
out["post_prob"] = {"columns_per_s": T / dt, "s": dt}
Tv = min(T, 20_000)                                # pure NumPy, 2.7e4 columns/s: a bounded slice
t0 = time.perf_counter()
path = ro.backtrack_viterbi(*ro.viterbi(a, b, pi, V[:Tv], order_nb))
dt = time.perf_counter() - t0
out["viterbi"] = {"columns_per_s": Tv / dt, "s": dt, "columns": Tv}
# the wrappers as the CLI calls them (each rebuilds `order`)
t0 = time.perf_counter()
llw = ro.loglik_wrapper(a, b, pi, [V])
out["loglik_wrapper_one_block_s"] = time.perf_counter() - t0
V_lst = [V[i * (T // cores):(i + 1) * (T // cores)] for i in range(cores)]
t0 = time.perf_counter()
ro.loglik_wrapper_par(a, b, pi, V_lst)
dt = time.perf_counter() - t0
out["loglik_wrapper_par"] = {"s": dt, "blocks": len(V_lst), "columns_per_s_incl_order_rebuild": T / dt}
# parity of this repo's oracle against the reference itself, on this host
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import hmm_oracle as ho                            # noqa: E402
out["oracle_vs_reference"] = {
    "loglik_rel": abs(ho.loglik_wrapper(a, b, pi, [V]) - float(ll)) / abs(float(ll)),
    "posterior_max_abs": float(np.abs(ho.post_prob_wrapper(a, b, pi, [V[:Tv]])[0] - post[:Tv]).max()) if Tv == T else None,
    "viterbi_equal": bool(np.array_equal(ho.viterbi_wrapper(a, b, pi, [V[:Tv]])[0], path))}
print(json.dumps(out, indent=1))
