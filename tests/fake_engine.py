"""Stand-in for ``itrails_b200.Engine`` that answers with the CPU oracle.

Used ONLY by the world-size-2 ``gloo`` tests (no GPU in the CPU suite): it lets the host
logic of the sharded path — LPT partition, all-reduce, gathering results back into input
order, part files spliced by rank 0 — run end to end.  The posterior CSV goes through
the real native writer (``itr_csv_posterior_host_ex``, host C++)."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import ctmc_oracle as co  # noqa: E402
import hmm_oracle as ho  # noqa: E402
from itrails_b200 import _lib as L  # noqa: E402


class FakeEngine:
    def __init__(self, device=0):
        self.device = device
        self.loads = 0
        self.V = []
        self.m = None
        self.sets = None
        self._post = None

    # data / model
    def load_blocks(self, V_lst):
        assert len(V_lst) > 0, "an empty share must never reach the engine"
        self.V = list(V_lst)
        self.loads += 1

    @property
    def n_blocks(self):
        return len(self.V)

    @property
    def n_columns(self):
        return int(sum(len(v) for v in self.V))

    def set_model(self, a, b, pi):
        self.m, self.sets = (a, b, pi), None
        self.K = np.asarray(a).shape[0]

    def build_model(self, params, n_ab, n_abc, cut_AB=None, cut_ABC=None, fetch=True):
        cab = "standard" if cut_AB is None else cut_AB
        cabc = "standard" if cut_ABC is None else cut_ABC
        outs = [co.trans_emiss_calc(*row, n_ab, n_abc, cab, cabc) for row in np.atleast_2d(params)]
        self.sets = [o[:3] for o in outs]
        self.m = self.sets[0]
        self.K = self.m[0].shape[0]
        if not fetch:
            return None, None, None, None
        hidden = np.array([outs[0][3][i] for i in range(self.K)], dtype=np.int32)
        return (np.stack([o[0] for o in outs]), np.stack([o[1] for o in outs]), np.stack([o[2] for o in outs]), hidden)

    # recursions
    def loglik(self):
        return np.array([ho.loglik_wrapper(*m, self.V) for m in (self.sets or [self.m])])

    def viterbi(self, log_a, log_E, omega0):
        return np.concatenate(ho.viterbi_wrapper(*self.m, self.V)).astype(np.uint8)

    def posterior(self, out=None, fetch=True):
        self._post = np.concatenate(ho.post_prob_wrapper(*self.m, self.V))
        return self._post if fetch else None

    def write_posterior_csv(self, path, positions=None, n_threads=0, block_ids=None, header=True):
        lib = L.load()
        off = np.concatenate([[0], np.cumsum([len(v) for v in self.V])]).astype(np.int64)
        nbytes = np.zeros(len(self.V), dtype=np.int64)
        post = np.ascontiguousarray(self._post)
        pos = None if positions is None else np.ascontiguousarray(positions, dtype=np.int64)
        ids = None if block_ids is None else np.ascontiguousarray(block_ids, dtype=np.int64)
        rc = lib.itr_csv_posterior_host_ex(os.fsencode(path), self.K, len(self.V), L.as_ptr(off, ctypes.c_int64),
                                           L.as_ptr(pos, ctypes.c_int64), L.as_ptr(post, ctypes.c_double),
                                           L.as_ptr(ids, ctypes.c_int64), 1 if header else 0,
                                           L.as_ptr(nbytes, ctypes.c_int64), 2)
        assert rc == 0
        return nbytes

    def close(self):
        pass
