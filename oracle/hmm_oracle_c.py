"""ctypes binding of oracle/hmm_oracle.c (TEST INFRASTRUCTURE ONLY — see that file).

Used where the NumPy restatement (oracle/hmm_oracle.py) is too slow: large-size
parity checks in tests/ and bench.py's timed CPU baseline."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libhmm_oracle.so")
_lib = None

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_bp = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        L.orc_loglik_blocks.argtypes = [ctypes.c_int, _dp, _dp, _dp, _ip, _ip,
                                        ctypes.c_int64, _dp, ctypes.c_int]
        L.orc_post_prob_blocks.argtypes = [ctypes.c_int, _dp, _dp, _dp, _ip, _ip,
                                           ctypes.c_int64, _dp, ctypes.c_int]
        L.orc_viterbi_blocks.argtypes = [ctypes.c_int, _dp, _dp, _dp, _ip, _ip,
                                         ctypes.c_int64, _bp, ctypes.c_int]
        for f in (L.orc_loglik_blocks, L.orc_post_prob_blocks, L.orc_viterbi_blocks):
            f.restype = None
        _lib = L
    return _lib


def _pack(V_lst):
    off = np.zeros(len(V_lst) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(v) for v in V_lst])
    V = np.ascontiguousarray(np.concatenate(V_lst).astype(np.int64))
    return V, off


def loglik_blocks(a, E, pi, V_lst, n_threads=1):
    V, off = _pack(V_lst)
    out = np.empty(len(V_lst))
    lib().orc_loglik_blocks(a.shape[0], np.ascontiguousarray(a), np.ascontiguousarray(E),
                            np.ascontiguousarray(pi), V, off, len(V_lst), out, n_threads)
    return out


def post_prob_blocks(a, E, pi, V_lst, n_threads=1):
    V, off = _pack(V_lst)
    K = a.shape[0]
    out = np.empty((len(V), K))
    lib().orc_post_prob_blocks(K, np.ascontiguousarray(a), np.ascontiguousarray(E),
                               np.ascontiguousarray(pi), V, off, len(V_lst), out, n_threads)
    return [out[off[i]:off[i + 1]] for i in range(len(V_lst))]


def viterbi_blocks(LA, LE, om0, V_lst, n_threads=1):
    V, off = _pack(V_lst)
    out = np.empty(len(V), dtype=np.uint8)
    lib().orc_viterbi_blocks(LA.shape[0], np.ascontiguousarray(LA), np.ascontiguousarray(LE),
                             np.ascontiguousarray(om0), V, off, len(V_lst), out, n_threads)
    return [out[off[i]:off[i + 1]] for i in range(len(V_lst))]
