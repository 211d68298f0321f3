"""Deterministic synthetic alignments of the shapes BASELINE.json names: columns are
sampled from the HMM itself (dwell-time sampling of the hidden path, then emitted
columns), and with probability ``p_n`` per column one random species is overwritten
by ``N`` so that symbols 256..624 occur."""
from __future__ import annotations

import numpy as np

from .read_data import _build_tables, get_obs_state_dct  # noqa: F401


class _Sampler:
    """Inverse-CDF tables of one model, shared by the blocks drawn from it."""

    def __init__(self, a, b, pi):
        a = np.asarray(a, dtype=np.float64)
        self.K = a.shape[0]
        self.leave = 1.0 - np.clip(np.diag(a), 0.0, 1.0 - 1e-12)
        off = a.copy()
        np.fill_diagonal(off, 0.0)
        off /= off.sum(1, keepdims=True)
        self.jump = [row.tolist() for row in np.cumsum(off, axis=1)]
        self.start = np.cumsum(pi / np.sum(pi)).tolist()
        emit = np.cumsum(b / b.sum(1, keepdims=True), axis=1)
        emit[:, -1] = 1.0
        # one sorted table for all states: entry 256 k + j = k + CDF_k(j)
        self.emit_flat = (np.arange(self.K)[:, None] + emit).ravel()


_SAMPLERS = {}


def _sampler(a, b, pi):
    key = (id(a), id(b), id(pi))
    hit = _SAMPLERS.get(key)
    if hit is None or hit[0] is not a:
        _SAMPLERS.clear()
        hit = _SAMPLERS[key] = (a, _Sampler(a, b, pi))
    return hit[1]


def sample_block(a, b, pi, T, rng, p_n=0.01, dtype=np.int64):
    """One block of ``T`` columns as symbol indices (like maf_parser's output): hidden
    path by dwell-time sampling (jump chain + geometric dwell times), columns emitted from
    the state's row of ``b``, then with probability ``p_n`` per column one random species
    is overwritten by ``N``."""
    import bisect
    from . import read_data as rd
    if rd._CODE_TO_INDEX is None:
        rd._build_tables()
    sm = _sampler(a, b, pi)
    K = sm.K
    zs, ds, total = [], [], 0
    z = min(bisect.bisect_left(sm.start, rng.random()), K - 1)
    while total < T:
        n = max(16, int((T - total) * float(sm.leave.mean()) * 1.5) + 16)
        u = rng.random(n).tolist()
        seg = [0] * n
        for k in range(n):
            seg[k] = z
            z = min(bisect.bisect_left(sm.jump[z], u[k]), K - 1)
        seg = np.array(seg, dtype=np.int64)
        d = rng.geometric(sm.leave[seg])
        zs.append(seg)
        ds.append(d)
        total += int(d.sum())
    seg, d = np.concatenate(zs), np.concatenate(ds)
    state = np.repeat(seg, d)[:T]
    # symbol = first j with CDF_state(j) > u, for all columns in one search of the flat table
    V = np.searchsorted(sm.emit_flat, state + rng.random(T), side="right") - 256 * state
    np.clip(V, 0, 255, out=V)
    hit = np.nonzero(rng.random(T) < p_n)[0]
    if len(hit):
        sp = rng.integers(0, 4, size=len(hit))
        v = V[hit]
        digits = np.stack([(v >> 6) & 3, (v >> 4) & 3, (v >> 2) & 3, v & 3], axis=1)
        digits[np.arange(len(hit)), sp] = 4
        code = ((digits[:, 0] * 5 + digits[:, 1]) * 5 + digits[:, 2]) * 5 + digits[:, 3]
        V[hit] = rd._CODE_TO_INDEX[code]
    return V if dtype == np.int64 else V.astype(dtype)


def block_lengths(n_blocks, total, rng, lo=50_000, hi=150_000):
    """``n_blocks`` lengths ~ uniform[lo, hi], rescaled so that they sum to ``total``."""
    raw = rng.uniform(lo, hi, size=n_blocks)
    lens = np.maximum(1, np.floor(raw * (total / raw.sum()))).astype(np.int64)
    lens[-1] += total - lens.sum()
    if lens[-1] < 1:
        raise ValueError("total too small for the number of blocks")
    return lens


def alignment(a, b, pi, lengths, seed, p_n=0.01):
    rng = np.random.default_rng(seed)
    return [sample_block(a, b, pi, int(T), rng, p_n) for T in lengths]


def alignment_blocks(a, b, pi, lengths, ids, seed, p_n=0.01, dtype=np.int64):
    """Blocks ``ids`` of the alignment whose block ``i`` has ``lengths[i]`` columns and its
    own random stream ``(seed, i)``: any process can generate any subset of the blocks and
    gets exactly the columns every other process would (sharded benchmarks generate only
    their share)."""
    return [sample_block(a, b, pi, int(lengths[int(i)]), np.random.default_rng([int(seed), int(i)]), p_n, dtype)
            for i in ids]


EXAMPLE_PARAMS = dict(mu=1e-8, N_AB=50000.0, N_ABC=50000.0, t_1=240000.0, t_2=40000.0,
                      t_upper=745069.3855, r=1e-8)


def example_model_args(n_int_ABC=3, p=None):
    """The nine scaled scalars of trans_emiss_calc for examples/example_config.yaml's
    starting values (scaling: workflow_optimize.py:369-405; t_out: optimizer.py:525-541)."""
    p = dict(EXAMPLE_PARAMS if p is None else p)
    mu = p["mu"]
    N_AB, N_ABC = p["N_AB"] * mu, p["N_ABC"] * mu
    t_1, t_2, t_upper, r = p["t_1"] * mu, p["t_2"] * mu, p["t_upper"] * mu, p["r"] / mu
    cut_last = -np.log1p(-(n_int_ABC - 1) / n_int_ABC)
    t_out = t_1 + t_2 + cut_last * N_ABC + t_upper + 2 * N_ABC
    return np.array([t_1, t_1, t_1 + t_2, t_2, t_upper, t_out, N_AB, N_ABC, r])


def write_maf(path, V_lst, species=("hg38", "panTro5", "gorGor5", "ponAbe2"), line_start=1000):
    """Write blocks of symbol indices as a MAF file (``+`` strand, one ``a`` block per
    array, ``N`` written as ``N``), so that ``maf_parser(path, species)`` returns
    ``V_lst`` again."""
    names = get_obs_state_dct()
    lut = np.array([[ord(c) for c in n] for n in names], dtype=np.uint8)     # 625 x 4
    with open(path, "wb") as fh:
        fh.write(b"##maf version=1 scoring=synthetic\n")
        pos = line_start
        for V in V_lst:
            cols = lut[np.asarray(V, dtype=np.int64)]                         # T x 4
            fh.write(b"a score=0\n")
            for k, sp in enumerate(species):
                seq = cols[:, k].tobytes()
                size = len(seq) - seq.count(b"-")
                fh.write(f"s {sp}.chr1 {pos} {size} + 250000000 ".encode() + seq + b"\n")
            fh.write(b"\n")
            pos += len(V)
