"""trans_emiss_calc — same signature and return value as the reference's
get_trans_emiss.py:8-170, computed by the batched CUDA model builder
(itr_build_model)."""
from __future__ import annotations

import numpy as np

from .engine_cache import get_engine
from .read_data import NUC


def _norm_cut(cut, n):
    if isinstance(cut, str):
        if cut != "standard":
            raise ValueError("cutpoints must be 'standard' or an array")
        return None
    cut = np.asarray(cut, dtype=np.float64)
    if cut.shape != (n + 1,):
        raise ValueError(f"expected {n + 1} cutpoints, got {cut.shape}")
    return cut


def trans_emiss_calc(t_A, t_B, t_C, t_2, t_upper, t_out, N_AB, N_ABC, r,
                     n_int_AB, n_int_ABC, cut_AB="standard", cut_ABC="standard"):
    """Returns ``(a, b, pi, hidden_names, observed_names)``: transition matrix
    (K,K), emission matrix (K,256), starting probabilities (K,), ``{index:
    (topology, i, j)}`` in sorted order and ``{index: 'AAAA'...}`` (nucleotide order
    A,C,T,G).  All inputs are in the reference's scaled units (times and population
    sizes multiplied by mu, recombination rate divided by mu)."""
    eng = get_engine()
    params = np.array([[t_A, t_B, t_C, t_2, t_upper, t_out, N_AB, N_ABC, r]], dtype=np.float64)
    a, b, pi, hidden = eng.build_model(params, int(n_int_AB), int(n_int_ABC),
                                       _norm_cut(cut_AB, int(n_int_AB)), _norm_cut(cut_ABC, int(n_int_ABC)))
    hidden_names = {i: (int(h[0]), int(h[1]), int(h[2])) for i, h in enumerate(hidden)}
    observed_names = {i: NUC[i >> 6] + NUC[(i >> 4) & 3] + NUC[(i >> 2) & 3] + NUC[i & 3]
                      for i in range(256)}
    return a[0], b[0], pi[0], hidden_names, observed_names
