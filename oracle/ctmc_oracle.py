"""CPU restatement (NumPy) of iTRAILS' model build: CTMC -> (a, b, pi).

TEST INFRASTRUCTURE ONLY.  Nothing under ``itrails_b200/`` may import this module.
Parity pinning: ``tests/test_oracle_model.py`` checks it against fixtures produced
by running the reference's ``trans_emiss_calc`` (``tests/golden/model_*.npz``,
made by ``tests/golden/make_golden.py``) for (n_int_AB, n_int_ABC) in
{(1,1), (2,1), (1,2), (2,2), (3,2), (1,3), (3,3)}.  At larger discretisations the
reference is intractable (hours per call); there this restatement IS the oracle and
parity is transitive.

What it restates (paths relative to /root/reference/src/itrails):

* ``trans_emiss_calc``             get_trans_emiss.py:8-170
* cutpoints                        cutpoints.py:5-65 (closed forms of the scipy ppf's)
* two-locus ARG state spaces       trans_mat.py:26-194, 269-286, 487-526
* ``combine_states``               combine_states.py:5-80
* ``get_joint_prob_mat``           get_joint_prob_mat.py:85-182
* ``run_markov_chain_AB``          run_markov_chain_AB.py:105-271
* ``run_markov_chain_ABC``         run_markov_chain_ABC.py:312-796
* ``vanloan`` / ``deepest_ti``     vanloan.py:392-425, deepest_ti.py:215-256
* JC69 emissions                   get_emission_prob_mat.py:47-92,120-397,484-698,701-1038

It is NOT a transliteration.  Three algebraic restatements (all checked against the
reference's numbers by the golden tests):

1. The reference sums Van Loan block-matrix exponentials over omega sub-paths
   (vanloan.py:392-425; run_markov_chain_ABC.py:61-115).  Because the per-locus
   coalescence class ("omega") only ever grows along a trajectory, that sum equals a
   sub-block of ``expm(dt * Q[S, S])`` where ``S`` is the set of states whose classes
   are allowed for the key.  For first-coalescence classes (x, y) the set
   ``S_xy = {0,x,7} x {0,y,7}`` (83 states) covers every key of an interval, so nine
   83x83 exponentials per interval replace all 203..1015-dim ones.
2. The last (infinite) interval's ``(-C^-1)[:n,-n:] @ A`` sums (deepest_ti.py:215-256)
   are absorption probabilities ``(-Q_TT)^-1 Q_T,term 1`` on the transient states of
   ``S_xy`` (those where some locus has not coalesced yet; <= 68 states).
3. The JC69 coalescent integrals (get_emission_prob_mat.py:47-92, 120-397) are
   evaluated by expanding each branch kernel ``1/4 + (delta-1/4) e^{-mu s}`` and
   integrating the resulting exponentials term by term, instead of the reference's
   expanded closed forms.
"""
from __future__ import annotations

import itertools
import math

import numpy as np

# ---------------------------------------------------------------------------
# cutpoints                                                   cutpoints.py:5-65
# ---------------------------------------------------------------------------


def cutpoints_AB(n_int_AB, t_AB, coal_AB):
    """Quantiles of an exponential(rate coal_AB) truncated to [0, t_AB]
    (cutpoints.py:5-26; scipy's truncexpon.ppf = -log1p(q*expm1(-b)))."""
    q = np.arange(n_int_AB + 1) / n_int_AB
    scale = 1 / coal_AB
    b = t_AB / scale
    with np.errstate(divide="ignore"):
        cut = -np.log1p(q * np.expm1(-b)) * scale
    cut[-1] = b * scale          # rv_continuous.ppf: q == 1 returns the upper end of the support
    return cut


def cutpoints_ABC(n_int_ABC, coal_ABC):
    """Quantiles of an exponential(rate coal_ABC); last one is +inf
    (cutpoints.py:29-45; expon.ppf = -log1p(-q))."""
    q = np.arange(n_int_ABC + 1) / n_int_ABC
    with np.errstate(divide="ignore"):
        return -np.log1p(-q) / coal_ABC


# ---------------------------------------------------------------------------
# dense matrix exponential (Higham 2005/2008 Pade-13 + scaling and squaring)
# expm.py:9-167 uses the same family; results agree to ~1e-15.
# ---------------------------------------------------------------------------
_PADE13 = (64764752532480000., 32382376266240000., 7771770303897600.,
           1187353796428800., 129060195264000., 10559470521600., 670442572800.,
           33522128640., 1323241920., 40840800., 960960., 16380., 182., 1.)


def expm(A):
    A = np.array(A, dtype=np.float64)
    n = A.shape[0]
    norm = np.abs(A).sum(axis=0).max() if n else 0.0
    s = 0
    if norm > 5.4:
        s = max(0, int(math.ceil(math.log2(norm / 5.4))))
        A = A / (2.0 ** s)
    c = _PADE13
    I = np.eye(n)
    A2 = A @ A
    A4 = A2 @ A2
    A6 = A2 @ A4
    U = A @ (A6 @ (c[13] * A6 + c[11] * A4 + c[9] * A2) + c[7] * A6 + c[5] * A4 + c[3] * A2 + c[1] * I)
    V = A6 @ (c[12] * A6 + c[10] * A4 + c[8] * A2) + c[6] * A6 + c[4] * A4 + c[2] * A2 + c[0] * I
    R = np.linalg.solve(V - U, V + U)
    for _ in range(s):
        R = R @ R
    return R


# ---------------------------------------------------------------------------
# state spaces                                 trans_mat.py:26-194, 269-286
# ---------------------------------------------------------------------------
class StateSpace:
    """Two-locus ancestral-recombination-graph state space for ``n`` species.

    A state is a set partition of the 2n lineage ends (positions 0..n-1 = left
    locus of species 1..n, n..2n-1 = right locus), stored as a restricted-growth
    label tuple (trans_mat.py:43-57 produces the same tuples, in another order —
    the order is immaterial to every result).  ``trans`` lists (from, to, kind):
    kind 1 = coalescence of two blocks (rate ``coal``), kind 2 = recombination of
    a block holding both left and right ends into its two halves (rate ``rho``)
    (trans_mat.py:74-194).  ``omega[i] = (l, r)``: bitmask over species whose
    left/right end shares a block with another end of the same locus
    (trans_mat.py:269-286)."""

    def __init__(self, n):
        self.n = n
        m = 2 * n
        states = []

        def rec(prefix, mx):
            if len(prefix) == m:
                states.append(tuple(prefix))
                return
            for lab in range(1, mx + 2):
                rec(prefix + [lab], max(mx, lab))

        rec([], 0)
        self.states = states
        self.index = {s: i for i, s in enumerate(states)}
        self.size = len(states)
        trans = []
        for i, s in enumerate(states):
            labs = sorted(set(s))
            for x, y in itertools.combinations(labs, 2):
                t = self._canon([x if v == y else v for v in s])
                trans.append((i, self.index[t], 1))
            for x in labs:
                pos = [k for k, v in enumerate(s) if v == x]
                if any(k < n for k in pos) and any(k >= n for k in pos):
                    new = max(s) + 1
                    t = self._canon([new if (v == x and k >= n) else v for k, v in enumerate(s)])
                    trans.append((i, self.index[t], 2))
        self.trans = np.array(trans, dtype=np.int64)
        om = []
        for s in states:
            pair = []
            for half in (s[:n], s[n:]):
                w = 0
                for k, v in enumerate(half):
                    if half.count(v) > 1:
                        w |= 1 << k
                pair.append(w)
            om.append(tuple(pair))
        self.omega = om
        self.omega_l = np.array([o[0] for o in om])
        self.omega_r = np.array([o[1] for o in om])

    @staticmethod
    def _canon(labels):
        m, out = {}, []
        for v in labels:
            if v not in m:
                m[v] = len(m) + 1
            out.append(m[v])
        return tuple(out)

    def generator(self, coal, rho):
        """trans_mat.py:487-508."""
        Q = np.zeros((self.size, self.size))
        f, t, k = self.trans.T
        Q[f, t] = np.where(k == 2, rho, coal)
        Q[np.arange(self.size), np.arange(self.size)] = -Q.sum(axis=1)
        return Q

    def mask(self, cl, cr):
        """Boolean mask of the states whose (left, right) classes lie in the
        given collections."""
        return np.isin(self.omega_l, list(cl)) & np.isin(self.omega_r, list(cr))


_SS = {}


def state_space(n):
    if n not in _SS:
        _SS[n] = StateSpace(n)
    return _SS[n]


def combine_states(ss1, ss2, ss12, v1, v2):
    """combine_states.py:5-80: the product distribution of two independent chains
    laid onto the merged chain (ends ordered left_1, left_2, right_1, right_2)."""
    out = np.zeros(ss12.size)
    n1, n2 = ss1.n, ss2.n
    for i1, s1 in enumerate(ss1.states):
        for i2, s2 in enumerate(ss2.states):
            off = max(s1)
            s2o = [v + off for v in s2]
            merged = list(s1[:n1]) + s2o[:n2] + list(s1[n1:]) + s2o[n2:]
            out[ss12.index[StateSpace._canon(merged)]] = v1[i1] * v2[i2]
    return out


# ---------------------------------------------------------------------------
# per-locus genealogy histories
# ---------------------------------------------------------------------------
# A history is (topology, t1, t2) exactly as in the reference's path keys
# (run_markov_chain_AB.py:139-146, run_markov_chain_ABC.py:368-392):
#   (-1,-1,-1)  nothing coalesced yet
#   (0, i, -1)  A,B coalesced in AB interval i           -> class 3
#   (k, s, -1)  first coalescence in ABC interval s, topology k in {1,2,3}
#               -> class 3 / 5 / 6
#   (k, s, u)   second coalescence in ABC interval u      -> class 7
NONE = (-1, -1, -1)
_CLASS_OF_TOPO = {0: 3, 1: 3, 2: 5, 3: 6}
_TOPO_OF_CLASS = {3: 1, 5: 2, 6: 3}


def hist_class(h):
    """helper_omegas.py:25-87 (one locus)."""
    if h[0] == -1:
        return 0
    return 7 if h[2] != -1 else _CLASS_OF_TOPO[h[0]]


def _succ_ABC(h, s):
    """Successor histories of one locus during ABC interval ``s`` and the
    first-coalescence class that constrains the move (None = unconstrained).
    run_markov_chain_ABC.py:368-392 plus the Van Loan key split
    vanloan.py:366-371."""
    if h[0] == -1:
        out = [(h, None)]
        for x, k in _TOPO_OF_CLASS.items():
            out.append(((k, s, -1), x))
            out.append(((k, s, s), x))
        return out
    if h[2] == -1:
        x = _CLASS_OF_TOPO[h[0]]
        return [(h, x), ((h[0], h[1], s), x)]
    return [(h, None)]


def joint_prob(t_A, t_B, t_AB, t_C, rho, coal_AB, coal_ABC, n_int_AB, n_int_ABC,
               cut_AB, cut_ABC):
    """get_joint_prob_mat.py:14-183 -> {(left history, right history): prob}.

    All recombination rates are equal and coal_A = coal_B = coal_C = coal_AB in
    the reference's only call site (get_trans_emiss.py:68-80)."""
    ss1, ss2, ss3 = state_space(1), state_space(2), state_space(3)
    Q1 = ss1.generator(coal_AB, rho)
    e0 = np.zeros(2)
    e0[ss1.index[(1, 1)]] = 1.0
    vA, vB, vC = (e0 @ expm(Q1 * t) for t in (t_A, t_B, t_C))

    # --- two-sequence chain, n_int_AB intervals      run_markov_chain_AB.py
    QAB = ss2.generator(coal_AB, rho)
    cls2 = {(l, r): ss2.mask([l], [r]) for l in (0, 3) for r in (0, 3)}
    cur = {(NONE, NONE): combine_states(ss1, ss1, ss2, vA, vB)}
    for s in range(n_int_AB):
        P = expm(QAB * (cut_AB[s + 1] - cut_AB[s]))
        nxt = {}
        for (hl, hr), v in cur.items():
            w = v @ P
            for hl2 in ([hl, (0, s, -1)] if hl[0] == -1 else [hl]):
                for hr2 in ([hr, (0, s, -1)] if hr[0] == -1 else [hr]):
                    nxt[(hl2, hr2)] = w * cls2[(hist_class(hl2), hist_class(hr2))]
        cur = nxt

    # --- merge with C                                get_joint_prob_mat.py:155-161
    cur = {k: combine_states(ss2, ss1, ss3, v, vC) for k, v in cur.items()}

    # --- three-sequence chain                        run_markov_chain_ABC.py
    Q = ss3.generator(coal_ABC, rho)
    firsts = (3, 5, 6)
    sets = {}
    for x in firsts:
        for y in firsts:
            idx = np.nonzero(ss3.mask([0, x, 7], [0, y, 7]))[0]
            sets[(x, y)] = idx
    cls_idx = {}
    for l in (0, 3, 5, 6, 7):
        for r in (0, 3, 5, 6, 7):
            cls_idx[(l, r)] = np.nonzero(ss3.mask([l], [r]))[0]

    def local(xy, cl):
        """positions of class ``cl``'s states inside S_xy"""
        return np.searchsorted(sets[xy], cls_idx[cl])

    for s in range(n_int_ABC - 1):
        dt = cut_ABC[s + 1] - cut_ABC[s]
        M = {xy: expm(Q[np.ix_(idx, idx)] * dt) for xy, idx in sets.items()}
        nxt = {}
        for (hl, hr), v in cur.items():
            c0 = (hist_class(hl), hist_class(hr))
            for hl2, x in _succ_ABC(hl, s):
                for hr2, y in _succ_ABC(hr, s):
                    xy = (x or 3, y or 3)
                    c1 = (hist_class(hl2), hist_class(hr2))
                    blk = M[xy][np.ix_(local(xy, c0), local(xy, c1))]
                    w = np.zeros(ss3.size)
                    w[cls_idx[c1]] = v[cls_idx[c0]] @ blk
                    nxt[(hl2, hr2)] = w
        cur = nxt

    # --- last interval (t -> inf)          run_markov_chain_ABC.py:519-795
    last = n_int_ABC - 1
    absorb = {}
    for (x, y), idx in sets.items():
        ol, orr = ss3.omega_l[idx], ss3.omega_r[idx]
        tr = (ol == 0) | (orr == 0)
        T, R = idx[tr], idx[~tr]
        w = np.linalg.solve(-Q[np.ix_(T, T)], Q[np.ix_(T, R)].sum(axis=1))
        full = np.zeros(ss3.size)
        full[T] = w
        absorb[(x, y)] = full

    def finals(h):
        if h[0] == -1:
            return [((k, last, last), x) for x, k in _TOPO_OF_CLASS.items()]
        if h[2] == -1:
            return [((h[0], h[1], last), _CLASS_OF_TOPO[h[0]])]
        return [(h, None)]

    out = {}
    for (hl, hr), v in cur.items():
        if hl[0] != -1 and hr[0] != -1:
            hl2 = hl if hl[2] != -1 else (hl[0], hl[1], last)
            hr2 = hr if hr[2] != -1 else (hr[0], hr[1], last)
            out[(hl2, hr2)] = v.sum()
            continue
        for hl2, x in finals(hl):
            for hr2, y in finals(hr):
                out[(hl2, hr2)] = v @ absorb[(x or 3, y or 3)]
    return out


# ---------------------------------------------------------------------------
# JC69 emissions                       get_emission_prob_mat.py:9-1038
# ---------------------------------------------------------------------------
_D4 = np.eye(4) - 0.25          # delta - 1/4
_Q4 = np.full((4, 4), 0.25)


def jc_branch(m):
    """4x4 JC69 transition matrix for accumulated rate*time ``m``
    (get_emission_prob_mat.py:9-44: expm(sum_i t_i Q_i), Q = mu/4 - mu I)."""
    return _Q4 + math.exp(-m) * _D4


def _int_exp(lam, t):
    """int_0^t e^{-lam u} du, stable near lam = 0."""
    return t if lam == 0.0 else -math.expm1(-lam * t) / lam


def single_coal_tensor(t, mu, k):
    """F[x1, x2, y] = sum_d int_0^t dens(u) P_{x1 d}(u) P_{x2 d}(u) P_{d y}(t-u) du
    with dens(u) = k e^{-k u} / (1 - e^{-k t}); restates
    get_emission_prob_mat.py:47-117 by term-wise integration."""
    norm = -math.expm1(-k * t)
    i1 = k * _int_exp(k + mu, t) / norm
    i2 = k * _int_exp(k + 2 * mu, t) / norm
    j0 = math.exp(-mu * t) * k * _int_exp(k - mu, t) / norm
    j1 = math.exp(-mu * t)
    j2 = j1 * i1
    D, Qm = _D4, _Q4
    es = np.einsum
    return (es("ad,bd,dy->aby", Qm, Qm, Qm)
            + i1 * (es("ad,bd,dy->aby", D, Qm, Qm) + es("ad,bd,dy->aby", Qm, D, Qm))
            + i2 * es("ad,bd,dy->aby", D, D, Qm)
            + j0 * es("ad,bd,dy->aby", Qm, Qm, D)
            + j1 * (es("ad,bd,dy->aby", D, Qm, D) + es("ad,bd,dy->aby", Qm, D, D))
            + j2 * es("ad,bd,dy->aby", D, D, D))


def double_coal_tensor(t, mu):
    """G[x1, x2, x3, y]: x1,x2 coalesce first (time u), their ancestor and x3
    second (time v), both inside an interval of length t (rates 3 then 1,
    conditioned on both events and on the topology), root evolves to y over t-v.
    Restates get_emission_prob_mat.py:120-424 by term-wise integration."""
    pboth = 1.0 + 0.5 * math.exp(-3 * t) - 1.5 * math.exp(-t)
    G = np.zeros((4, 4, 4, 4))
    mats = (_Q4, _D4)
    for na, nb, ng, nd, ne in itertools.product((0, 1), repeat=5):
        p = 2.0 + mu * (na + nb - ng)
        q = 1.0 + mu * (ng + nd - ne)
        integ = (_int_exp(p + q, t) - math.exp(-q * t) * _int_exp(p, t)) / q
        coef = 3.0 / pboth * math.exp(-mu * ne * t) * integ
        G += coef * np.einsum("ae,be,ef,cf,fd->abcd", mats[na], mats[nb], mats[ng],
                              mats[nd], mats[ne])
    return G


def emission_table(t_A, t_B, t_AB, t_C, t_upper, t_out, coal_AB, coal_ABC, mu,
                   n_int_AB, n_int_ABC, cut_AB, cut_ABC):
    """get_emission_prob_mat.py:701-1038 -> (states, b) with ``b[k]`` a 256-vector
    in observed order 64a+16b+4c+d (A,C,T,G = 0..3).  All branch mutation rates
    are ``mu`` (get_trans_emiss.py:84-89)."""
    n = n_int_ABC
    states, rows = [], []

    def width(j):       # get_emission_prob_mat.py:818-820
        return cut_ABC[j + 1] - cut_ABC[j] if j != n - 1 else t_upper

    def above(j):       # get_emission_prob_mat.py:822-826
        return t_upper + cut_ABC[n - 1] - cut_ABC[j + 1] if j != n - 1 else 0.0

    def single(Pa, Pb, F1, Pab, F2, Pc, Pd):   # :585-606
        return np.einsum("ai,jb,ijk,kl,lmn,mc,nd->abcd", Pa, Pb, F1, Pab, F2, Pc, Pd) / 4

    def double(Pa, Pb, Pc, G, Pd):             # :681-697
        return np.einsum("ai,jb,kc,ijkn,nd->abcd", Pa, Pb, Pc, G, Pd) / 4

    for i in range(n):
        for j in range(i + 1, n):
            Pa = jc_branch(mu * (t_A + t_AB + cut_ABC[i]))
            Pb = jc_branch(mu * (t_B + t_AB + cut_ABC[i]))
            Pc = jc_branch(mu * (t_C + cut_ABC[i]))
            Pab = jc_branch(mu * (cut_ABC[j] - cut_ABC[i + 1]))
            F1 = single_coal_tensor(cut_ABC[i + 1] - cut_ABC[i], mu, coal_ABC)
            F2 = single_coal_tensor(width(j), mu, coal_ABC)
            Pd = jc_branch(mu * (t_out + above(j)))
            states.append((1, i, j))
            rows.append(single(Pa, Pb, F1, Pab, F2, Pc, Pd))
            states.append((2, i, j))
            rows.append(single(Pa, Pc, F1, Pab, F2, Pb, Pd).transpose(0, 2, 1, 3))
            states.append((3, i, j))
            rows.append(single(Pb, Pc, F1, Pab, F2, Pa, Pd).transpose(2, 0, 1, 3))
    for i in range(n):
        Pa = jc_branch(mu * (t_A + t_AB + cut_ABC[i]))
        Pb = jc_branch(mu * (t_B + t_AB + cut_ABC[i]))
        Pc = jc_branch(mu * (t_C + cut_ABC[i]))
        G = double_coal_tensor(width(i), mu)
        Pd = jc_branch(mu * (t_out + above(i)))
        states.append((1, i, i))
        rows.append(double(Pa, Pb, Pc, G, Pd))
        states.append((2, i, i))
        rows.append(double(Pa, Pc, Pb, G, Pd).transpose(0, 2, 1, 3))
        states.append((3, i, i))
        rows.append(double(Pb, Pc, Pa, G, Pd).transpose(2, 0, 1, 3))
    for i in range(n_int_AB):
        for j in range(n):
            Pa = jc_branch(mu * (t_A + cut_AB[i]))
            Pb = jc_branch(mu * (t_B + cut_AB[i]))
            Pc = jc_branch(mu * (t_C + cut_ABC[j]))
            Pab = jc_branch(mu * (t_AB - cut_AB[i + 1] + cut_ABC[j]))
            F1 = single_coal_tensor(cut_AB[i + 1] - cut_AB[i], mu, coal_AB)
            F2 = single_coal_tensor(width(j), mu, coal_ABC)
            Pd = jc_branch(mu * (t_out + above(j)))
            states.append((0, i, j))
            rows.append(single(Pa, Pb, F1, Pab, F2, Pc, Pd))
    return states, np.array([r.reshape(256) for r in rows])


# ---------------------------------------------------------------------------
# top level                                         get_trans_emiss.py:8-170
# ---------------------------------------------------------------------------
def trans_emiss_calc(t_A, t_B, t_C, t_2, t_upper, t_out, N_AB, N_ABC, r,
                     n_int_AB, n_int_ABC, cut_AB="standard", cut_ABC="standard"):
    """Same signature and return value as the reference's ``trans_emiss_calc``:
    ``(a, b, pi, hidden_names, observed_names)``."""
    N_ref = N_ABC
    t_A, t_B, t_AB, t_C = t_A / N_ref, t_B / N_ref, t_2 / N_ref, t_C / N_ref
    t_upper, t_out = t_upper / N_ref, t_out / N_ref
    rho = N_ref * r
    coal_AB = N_ref / N_AB
    coal_ABC = N_ref / N_ABC
    mu = N_ref * (4 / 3)
    if isinstance(cut_AB, str):
        cut_AB = cutpoints_AB(n_int_AB, t_AB, coal_AB)
    if isinstance(cut_ABC, str):
        cut_ABC = cutpoints_ABC(n_int_ABC, coal_ABC)
    cut_AB = np.asarray(cut_AB, dtype=np.float64)
    cut_ABC = np.asarray(cut_ABC, dtype=np.float64)

    joint = joint_prob(t_A, t_B, t_AB, t_C, rho, coal_AB, coal_ABC, n_int_AB,
                       n_int_ABC, cut_AB, cut_ABC)
    states, b = emission_table(t_A, t_B, t_AB, t_C, t_upper, t_out, coal_AB,
                               coal_ABC, mu, n_int_AB, n_int_ABC, cut_AB, cut_ABC)
    order = sorted(range(len(states)), key=lambda i: states[i])
    hidden = [states[i] for i in order]
    b = b[order]
    index = {h: i for i, h in enumerate(hidden)}
    K = len(hidden)
    J = np.zeros((K, K))
    for (hl, hr), p in joint.items():
        J[index[hl], index[hr]] = p
    pi = J.sum(axis=1)
    a = J / pi[:, None]
    nuc = "ACTG"
    observed = {i: nuc[i >> 6] + nuc[(i >> 4) & 3] + nuc[(i >> 2) & 3] + nuc[i & 3]
                for i in range(256)}
    return a, b, pi, dict(enumerate(hidden)), observed
