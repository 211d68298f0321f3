"""CPU restatement of iTRAILS' HMM recursions and observed-symbol alphabet.

TEST INFRASTRUCTURE ONLY.  Nothing under ``itrails_b200/`` may import this module;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg do,
and only as the checker.  Parity pinning: checked against fixtures produced by
running the reference itself (``tests/golden/make_golden.py`` ->
``tests/golden/symbols.npz`` / ``recursions_*.npz``) in ``tests/test_oracle_hmm.py``.

Each function cites the reference lines it restates (paths relative to
``/root/reference/src/itrails``).  The arithmetic follows the reference step by
step (log-space forward/backward with a running maximum; max-plus Viterbi with
``argmax`` = first maximum) so the numbers agree to the last few ulps, and the
Viterbi path agrees exactly.
"""
from __future__ import annotations

import numpy as np

NUC = "ACTG"  # reference nucleotide order, read_data.py:13 (NOT "ACGT")
N_SYMBOLS = 625
N_PLAIN = 256


# ---------------------------------------------------------------------------
# observed symbols                                   read_data.py:6-24, 46-67
# ---------------------------------------------------------------------------
def obs_state_names():
    """The 625 four-letter column strings in the reference's index order.

    read_data.py:6-24: first the 256 N-free strings with A,C,T,G nested loops
    (index = 64a+16b+4c+d), then every string over A,C,T,G,N not yet listed, in
    nested-loop order."""
    plain = [a + b + c + d for a in NUC for b in NUC for c in NUC for d in NUC]
    seen = set(plain)
    ext = NUC + "N"
    rest = [a + b + c + d for a in ext for b in ext for c in ext for d in ext
            if a + b + c + d not in seen]
    return plain + rest


def code_to_index_table():
    """Base-5 column code (digits A,C,T,G,N = 0..4, first species most significant) ->
    index in ``obs_state_names()``."""
    names = obs_state_names()
    ext = NUC + "N"
    tab = np.zeros(625, dtype=np.int64)
    for i, s in enumerate(names):
        code = 0
        for ch in s:
            code = code * 5 + ext.index(ch)
        tab[code] = i
    return tab


def order_lists():
    """``order[s]``: the N-free symbol indices symbol ``s`` marginalises over.

    read_data.py:46-67 expands the first ``N`` into A,C,T,G recursively, which
    yields the compatible N-free indices in ascending order."""
    names = obs_state_names()
    out = []
    for s in names:
        idx = [0]
        for ch in s:
            if ch == "N":
                idx = [4 * i + k for i in idx for k in range(4)]
            else:
                k = NUC.index(ch)
                idx = [4 * i + k for i in idx]
        out.append(np.array(idx, dtype=np.int64))
    return out


def emission_table(b, order=None):
    """E[:, s] = b[:, order[s]].sum(axis=1) for all 625 symbols.

    Uses the reference's exact NumPy expression (optimizer.py:182,186,329) so the
    pairwise-summation rounding is identical."""
    order = order_lists() if order is None else order
    K = b.shape[0]
    E = np.empty((K, N_SYMBOLS))
    for s in range(N_SYMBOLS):
        E[:, s] = b[:, order[s]].sum(axis=1)
    return E


# ---------------------------------------------------------------------------
# forward / backward / posterior                         optimizer.py:146-238
# ---------------------------------------------------------------------------
def forward(a, b, pi, V, E=None):
    """optimizer.py:166-188."""
    E = emission_table(b) if E is None else E
    T, K = len(V), a.shape[0]
    alpha = np.zeros((T, K))
    with np.errstate(divide="ignore"):
        alpha[0] = np.log(pi * E[:, V[0]])
        for t in range(1, T):
            x = alpha[t - 1].max()
            alpha[t] = np.log((np.exp(alpha[t - 1] - x) @ a) * E[:, V[t]]) + x
    return alpha


def forward_loglik(a, b, pi, V, E=None):
    """optimizer.py:146-162."""
    alpha = forward(a, b, pi, V, E)
    x = alpha[-1].max()
    return np.log(np.exp(alpha[-1] - x).sum()) + x


def backward(a, b, V, E=None):
    """optimizer.py:192-213.  NOTE the reference multiplies the ROW vector by
    ``a`` (``(exp(beta)·e) @ a``), i.e. it uses a-transposed relative to the
    textbook recursion; restated as is."""
    E = emission_table(b) if E is None else E
    T, K = len(V), a.shape[0]
    beta = np.zeros((T, K))
    with np.errstate(divide="ignore"):
        for t in range(T - 2, -1, -1):
            x = beta[t + 1].max()
            beta[t] = np.log((np.exp(beta[t + 1] - x) * E[:, V[t + 1]]) @ a) + x
    return beta


def post_prob(a, b, pi, V, E=None):
    """optimizer.py:216-238."""
    E = emission_table(b) if E is None else E
    s = forward(a, b, pi, V, E) + backward(a, b, V, E)
    m = s.max(1).reshape(-1, 1)
    w = np.exp(s - m)
    return w / w.sum(1).reshape(-1, 1)


def loglik_wrapper(a, b, pi, V_lst):
    """optimizer.py:93-116."""
    E = emission_table(b)
    acc = 0
    for V in V_lst:
        acc += forward_loglik(a, b, pi, V, E)
    return acc


def post_prob_wrapper(a, b, pi, V_lst):
    """optimizer.py:241-262."""
    E = emission_table(b)
    return [post_prob(a, b, pi, V, E) for V in V_lst]


# ---------------------------------------------------------------------------
# Viterbi                                                optimizer.py:305-377
# ---------------------------------------------------------------------------
def viterbi_tables(a, b, pi, V_lst):
    """Host-side tables whose exact FP64 values decide the Viterbi path:
    ``LA = log a``, ``LE = log E`` and per block ``omega0 = log(pi * E[:, V0])``
    (optimizer.py:323, 327-330)."""
    E = emission_table(b)
    with np.errstate(divide="ignore"):
        LA = np.log(a)
        LE = np.log(E)
        om0 = np.array([np.log(pi * E[:, V[0]]) for V in V_lst])
    return LA, LE, om0


def viterbi(a, b, pi, V, E=None):
    """optimizer.py:305-333 — returns (omega, prev) like the reference."""
    E = emission_table(b) if E is None else E
    T, K = len(V), a.shape[0]
    omega = np.zeros((T, K))
    prev = np.zeros((T - 1, K))
    with np.errstate(divide="ignore"):
        omega[0] = np.log(pi * E[:, V[0]])
        LA = np.log(a)
        for t in range(1, T):
            M = omega[t - 1][:, np.newaxis] + LA + np.log(E[:, V[t]])
            prev[t - 1] = np.argmax(M, axis=0)
            omega[t] = np.max(M, axis=0)
    return omega, prev


def backtrack_viterbi(omega, prev):
    """optimizer.py:336-354 — float64 state path, first maximum at the end."""
    T = omega.shape[0]
    S = np.zeros(T)
    last = int(np.argmax(omega[T - 1]))
    S[T - 1] = last
    for i in range(T - 2, -1, -1):
        last = int(prev[i, last])
        S[i] = last
    return S


def viterbi_wrapper(a, b, pi, V_lst):
    """optimizer.py:357-377."""
    E = emission_table(b)
    return [backtrack_viterbi(*viterbi(a, b, pi, V, E)) for V in V_lst]


# ---------------------------------------------------------------------------
# synthetic data: columns sampled from the HMM itself
# ---------------------------------------------------------------------------
def sample_block(a, b, pi, T, rng, p_n=0.01):
    """Dwell-time sampling of a hidden path + emitted columns, then with
    probability ``p_n`` per column one random species is overwritten by ``N``
    (symbols 256..624).  Returns int64 symbols like ``maf_parser``."""
    K = a.shape[0]
    names = obs_state_names()
    lookup = {s: i for i, s in enumerate(names)}
    stay = np.clip(np.diag(a), 0.0, 1.0 - 1e-12)
    off = a.copy()
    np.fill_diagonal(off, 0.0)
    off /= off.sum(1, keepdims=True)
    bc = np.cumsum(b / b.sum(1, keepdims=True), axis=1)
    z = rng.choice(K, p=pi / pi.sum())
    V = np.empty(T, dtype=np.int64)
    t = 0
    while t < T:
        d = int(rng.geometric(1.0 - stay[z]))
        d = min(d, T - t)
        u = rng.random(d)
        V[t:t + d] = np.minimum(np.searchsorted(bc[z], u), 255)
        t += d
        z = rng.choice(K, p=off[z])
    hit = np.nonzero(rng.random(T) < p_n)[0]
    sp = rng.integers(0, 4, size=len(hit))
    for t, k in zip(hit, sp):
        s = list(names[V[t]])
        s[k] = "N"
        V[t] = lookup["".join(s)]
    return V
