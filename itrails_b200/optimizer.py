"""HMM wrappers and optimiser driver — same names, arguments and return types as the
reference's optimizer.py, with the numerics on the GPU:

    loglik_wrapper / loglik_wrapper_par   optimizer.py:40-116
    post_prob_wrapper                     optimizer.py:241-262
    viterbi_wrapper                       optimizer.py:357-377
    optimization_wrapper / optimizer      optimizer.py:396-637
    write_list                            optimizer.py:380-393

With more than one GPU (one process per GPU under ``torch.distributed``, and / or
several GPUs driven by this process — ``itrails_b200.ngpu``) the blocks are
LPT-partitioned over the GPUs; the log-likelihood wrappers all-reduce the scalar, and
the decoders gather their results so that EVERY wrapper returns one entry per ``V_lst``
block in input order on every rank, exactly like the single-GPU call
(``itrails_b200.distributed``).
"""
from __future__ import annotations

import os
import time

import numpy as np

from . import distributed as dist_
from . import ngpu
from .cutpoints import cutpoints_ABC
from .engine_cache import ensure_blocks, get_engine  # noqa: F401  (get_engine: re-exported for callers)
from .read_data import order_lists
from .yaml_helpers import update_best_model


# ---------------------------------------------------------------------------
# host-side tables
# ---------------------------------------------------------------------------
def emission_table(b):
    """E[:, s] = b[:, order[s]].sum(axis=1) for the 625 observed symbols — the
    reference's own NumPy expression (optimizer.py:182, 329), so the rounding of the
    table Viterbi's decisions depend on is identical."""
    order = order_lists()
    b = np.asarray(b, dtype=np.float64)
    E = np.empty((b.shape[0], 625))
    for s in range(625):
        E[:, s] = b[:, order[s]].sum(axis=1)
    return E


def viterbi_tables(a, b, pi, V_lst):
    """log a, log E and per-block omega_0 = log(pi * e(V_0)) computed with NumPy
    exactly as optimizer.py:323-330 does."""
    E = emission_table(b)
    pi = np.asarray(pi, dtype=np.float64)
    with np.errstate(divide="ignore"):
        log_a = np.log(np.asarray(a, dtype=np.float64))
        log_E = np.log(E)
        first = np.array([int(V[0]) for V in V_lst], dtype=np.int64)
        omega0 = np.log(pi[None, :] * E[:, first].T)
    return log_a, log_E, np.ascontiguousarray(omega0)


# ---------------------------------------------------------------------------
# sharding: which blocks live on which GPU, and putting results back in input order
# ---------------------------------------------------------------------------
class _Shard:
    """One GPU's share of ``V_lst``: ``ids`` are the global block indices (ascending)."""
    __slots__ = ("device", "local", "ids", "eng")

    def __init__(self, device, local, ids):
        self.device, self.local, self.ids, self.eng = device, local, ids, None


class _Plan:
    __slots__ = ("V_lst", "shards", "parts", "lengths", "n_local", "world", "rank")


_PLAN_CACHE = {}


def _plan(V_lst):
    """LPT partition of ``V_lst`` over every GPU of the job (cached per list object), with
    this process's shares resident in HBM."""
    hit = _PLAN_CACHE.get(id(V_lst))
    if hit is not None and hit.V_lst is V_lst and len(hit.lengths) == len(V_lst):
        plan = hit
    else:
        _PLAN_CACHE.clear()
        plan = _Plan()
        plan.V_lst = V_lst
        plan.rank, plan.world = dist_.rank_world()
        devices = ngpu.local_devices()
        plan.n_local = len(devices)
        plan.lengths = np.array([len(v) for v in V_lst], dtype=np.int64)
        n_parts = plan.world * plan.n_local
        if n_parts == 1:
            plan.parts = [np.arange(len(V_lst), dtype=np.int64)]
            plan.shards = [_Shard(devices[0], V_lst, plan.parts[0])]
        else:
            plan.parts = dist_.lpt_partition(plan.lengths, n_parts)
            plan.shards = []
            for g, dev in enumerate(devices):
                ids = plan.parts[plan.rank * plan.n_local + g]
                plan.shards.append(_Shard(dev, [V_lst[i] for i in ids], ids))
        _PLAN_CACHE[id(V_lst)] = plan
    for sh in plan.shards:              # (an empty share is legal: fewer blocks than GPUs)
        sh.eng = ensure_blocks(sh.local, sh.device) if len(sh.ids) else None
    return plan


def _each(plan, fn):
    """fn(shard) for every non-empty local shard; concurrently when this process drives
    several GPUs (the C ABI calls release the GIL).  Returns results in shard order
    (None for empty shards)."""
    live = [sh for sh in plan.shards if sh.eng is not None]
    if len(live) <= 1:
        res = {id(sh): fn(sh) for sh in live}
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(len(live)) as pool:
            res = dict(zip((id(sh) for sh in live), pool.map(fn, live)))
    return [res.get(id(sh)) for sh in plan.shards]


def _global_order(plan, local_flat, dtype, trailing=()):
    """Per-shard flat results (columns of the shard's blocks back to back, None for empty
    shards) -> list with one array per block of ``V_lst`` in input order, on every rank."""
    n = len(plan.lengths)
    if plan.world * plan.n_local == 1:
        off = np.concatenate([[0], np.cumsum(plan.lengths)])
        flat = local_flat[0]
        return [flat[off[i]:off[i + 1]] for i in range(n)]
    mine = [f if f is not None else np.empty((0,) + tuple(trailing), dtype=dtype) for f in local_flat]
    per_rank = [mine]
    if plan.world > 1:
        # one contribution per rank: its shards' results back to back
        counts = [int(sum(plan.lengths[plan.parts[r * plan.n_local + g]].sum() for g in range(plan.n_local)))
                  for r in range(plan.world)]
        cat = np.concatenate(mine) if len(mine) > 1 else mine[0]
        gathered = dist_.allgather_ragged(cat, counts)
        per_rank = []
        for r in range(plan.world):
            cuts = np.cumsum([int(plan.lengths[plan.parts[r * plan.n_local + g]].sum()) for g in range(plan.n_local)])[:-1]
            per_rank.append(np.split(gathered[r], cuts) if plan.n_local > 1 else [gathered[r]])
    out = [None] * n
    for r, shards in enumerate(per_rank):
        for g, flat in enumerate(shards):
            ids = plan.parts[(r if plan.world > 1 else plan.rank) * plan.n_local + g]
            off = np.concatenate([[0], np.cumsum(plan.lengths[ids])])
            for k, i in enumerate(ids):
                out[int(i)] = flat[off[k]:off[k + 1]]
    return out


def _sum_loglik(plan, partials):
    """Sum of per-shard log-likelihood vectors over local shards, then over ranks."""
    vecs = [p for p in partials if p is not None]
    n_sets = len(vecs[0]) if vecs else None
    if plan.world > 1 and any(len(p) == 0 for p in plan.parts):
        # fewer blocks than GPUs: a rank without blocks still takes part in the all-reduce,
        # with the vector length the others use (every rank sees the same partition, so
        # either all of them make this extra exchange or none does)
        n_sets = int(dist_.allreduce_max(-1 if n_sets is None else n_sets))
    total = np.sum(vecs, axis=0) if vecs else np.zeros(n_sets)
    if plan.world > 1:
        dev = plan.shards[0].device
        total = dist_.allreduce_sum(total, dev)
    return np.asarray(total, dtype=np.float64)


# ---------------------------------------------------------------------------
# wrappers
# ---------------------------------------------------------------------------
def loglik_wrapper(a, b, pi, V_lst):
    """Sum over blocks of the forward log-likelihood (optimizer.py:93-116)."""
    plan = _plan(V_lst)

    def one(sh):
        sh.eng.set_model(a, b, pi)
        return sh.eng.loglik()

    total = _sum_loglik(plan, _each(plan, one))
    return float(total[0])      # a Python float, as the reference's numba forward_loglik returns


def loglik_wrapper_par(a, b, pi, V_lst):
    """optimizer.py:40-65 — the reference farms blocks to joblib workers; here every
    block is a warp-level chain on the GPU, so this is ``loglik_wrapper``."""
    return loglik_wrapper(a, b, pi, V_lst)


def post_prob_wrapper(a, b, pi, V_lst):
    """List of (T, K) float64 posterior matrices, one per block of ``V_lst`` in input
    order (optimizer.py:241-262).  On several GPUs every rank receives the full list
    (the shards are exchanged after decoding); for results that do not fit in host
    memory use ``post_prob_to_csv``."""
    plan = _plan(V_lst)
    K = np.asarray(a).shape[0]

    def one(sh):
        sh.eng.set_model(a, b, pi)
        return sh.eng.posterior()

    return _global_order(plan, _each(plan, one), np.float64, trailing=(K,))


def _csv_header(K):
    return ("alignment_block_idx,position_idx" + "".join(f",prob_state_{i}" for i in range(K)) + "\r\n").encode()


def post_prob_to_csv(a, b, pi, V_lst, output_file, ref_coordinates=None, n_threads=0):
    """Posterior decoding written straight to ``output_file`` in the reference's format
    (workflow_posterior.py:697-716) without materialising the list of (T, K) matrices on
    the host: the result stays in HBM and the native writer streams it block by block.
    Byte-identical to ``csv.writer`` over ``post_prob_wrapper``'s result.  On several
    GPUs every GPU writes the rows of its own blocks — with their GLOBAL block index and
    coordinates — to a part file, and rank 0 splices the parts in input block order."""
    plan = _plan(V_lst)
    if ref_coordinates is not None and [len(c) for c in ref_coordinates] != [int(n) for n in plan.lengths]:
        raise IndexError("reference coordinates do not match the alignment blocks")
    single = plan.world * plan.n_local == 1

    def positions_of(ids):
        if ref_coordinates is None:
            return None
        return np.concatenate([np.asarray(ref_coordinates[int(i)], dtype=np.int64) for i in ids])

    def one(sh):
        sh.eng.set_model(a, b, pi)
        sh.eng.posterior(fetch=False)
        if single:
            sh.eng.write_posterior_csv(output_file, positions_of(sh.ids), n_threads)
            return None
        part = f"{output_file}.part{plan.rank * plan.n_local + plan.shards.index(sh)}"
        return part, sh.eng.write_posterior_csv(part, positions_of(sh.ids), n_threads, block_ids=sh.ids, header=False)

    res = _each(plan, one)
    if single:
        return
    # (part file, bytes per block) of every part of the job, in part order
    mine = [(r[0], r[1].tolist()) if r is not None else (None, []) for r in res]
    parts_info = [x for per_rank in dist_.allgather_object(mine) for x in per_rank]
    if plan.rank == 0:
        K = np.asarray(a).shape[0]
        handles = [open(pth, "rb") if pth else None for pth, _ in parts_info]
        cursor = [0] * len(parts_info)          # next block of each part (parts hold ascending ids)
        owner = np.empty(len(plan.lengths), dtype=np.int64)
        for q, ids in enumerate(plan.parts):
            owner[ids] = q
        with open(output_file, "wb") as out:
            out.write(_csv_header(K))
            for i in range(len(plan.lengths)):
                q = int(owner[i])
                left = parts_info[q][1][cursor[q]]
                cursor[q] += 1
                while left > 0:
                    buf = handles[q].read(min(left, 1 << 24))
                    if not buf:
                        raise OSError(f"part file {parts_info[q][0]} is shorter than its index")
                    out.write(buf)
                    left -= len(buf)
        for h, (pth, _) in zip(handles, parts_info):
            if h is not None:
                h.close()
                os.remove(pth)
    dist_.barrier()


def viterbi_wrapper(a, b, pi, V_lst):
    """List of float64 state paths, one per block of ``V_lst`` in input order
    (optimizer.py:357-377; the reference returns float arrays, e.g. ``14.0``)."""
    plan = _plan(V_lst)

    def one(sh):
        sh.eng.set_model(a, b, pi)
        return sh.eng.viterbi(*viterbi_tables(a, b, pi, sh.local))

    return [p.astype(np.float64) for p in _global_order(plan, _each(plan, one), np.uint8)]


def _objective_batch(rows, n_int_AB, n_int_ABC, V_lst):
    """Summed log-likelihood of every parameter row: batched model build + multi-set
    forward sweep on every GPU's share, then the scalar reduction."""
    plan = _plan(V_lst)
    rows = np.ascontiguousarray(rows, dtype=np.float64)

    def one(sh):
        sh.eng.build_model(rows, n_int_AB, n_int_ABC, fetch=False)
        return sh.eng.loglik()

    return _sum_loglik(plan, _each(plan, one))


# ---------------------------------------------------------------------------
# optimiser driver
# ---------------------------------------------------------------------------
def write_list(lst, res_name):
    """Append one comma-separated line (optimizer.py:380-393)."""
    with open(f"{res_name}", "a") as fh:
        fh.write(",".join(str(x) for x in lst) + "\n")


def derive_times(d, case):
    """Fill in t_A, t_B, t_C and t_out from the user's time parametrisation
    ("case"), following optimizer.py:419-541.  ``d`` is modified in place and
    returned; a user-fixed ``t_out`` is kept."""
    n = d["n_int_ABC"]
    tail = cutpoints_ABC(n, 1)[n - 1] * d["N_ABC"] + d["t_upper"] + 2 * d["N_ABC"]
    case = frozenset(case)
    if "t_1" in case:
        t_1 = d.pop("t_1")
        d.setdefault("t_A", t_1)
        d.setdefault("t_B", t_1)
        d["t_C"] = d["t_C"] if "t_C" in case else t_1 + d["t_2"]
        t_out = t_1 + d["t_2"] + tail
    else:
        if case == frozenset(["t_A", "t_B"]):
            d["t_C"] = (d["t_A"] + d["t_B"]) / 2 + d["t_2"]
        elif case == frozenset(["t_A", "t_C"]):
            d["t_B"] = (d["t_A"] + d["t_C"] - d["t_2"]) / 2
        elif case == frozenset(["t_B", "t_C"]):
            d["t_A"] = (d["t_B"] + d["t_C"] - d["t_2"]) / 2
        elif case != frozenset(["t_A", "t_B", "t_C"]):
            raise ValueError(f"Invalid combination of time values: {set(case)}")
        t_out = (((d["t_A"] + d["t_B"]) / 2 + d["t_2"]) + d["t_C"]) / 2 + tail
    d.setdefault("t_out", t_out)
    return d


def model_args(d):
    return (d["t_A"], d["t_B"], d["t_C"], d["t_2"], d["t_upper"], d["t_out"],
            d["N_AB"], d["N_ABC"], d["r"])


def optimization_wrapper(arg_lst, optimized_params, case, d, V_lst, res_name, info):
    """Objective (optimizer.py:396-583): parameters -> model build -> log-likelihood,
    appended to ``<prefix>.optimization_history.csv``; ``<prefix>.best_model.yaml`` is
    rewritten on improvement; returns ``-loglik``."""
    output_dir, output_prefix = os.path.split(res_name)
    best_model_yaml = os.path.join(output_dir, f"{output_prefix}.best_model.yaml")
    d_copy = d.copy()
    for i, param in enumerate(optimized_params):
        d_copy[param] = arg_lst[i]
    derive_times(d_copy, case)
    # trans_emiss_calc + loglik_wrapper of the reference (optimizer.py:543-567), without the
    # round trip of (a, b, pi) through the host: the builder leaves the model installed on
    # the device next to this rank's resident blocks.
    total = _objective_batch(np.array([model_args(d_copy)]), d_copy["n_int_AB"], d_copy["n_int_ABC"], V_lst)
    loglik = float(total[0])
    rank, _world = dist_.rank_world()
    if rank == 0:
        write_list([info["Nfeval"]] + np.asarray(arg_lst).tolist() + [loglik, time.time() - info["time"]],
                   os.path.join(output_dir, f"{output_prefix}.optimization_history.csv"))
        update_best_model(best_model_yaml, optimized_params, arg_lst, loglik, info["Nfeval"])
    info["Nfeval"] += 1
    return -loglik


def loglik_sweep(arg_sets, optimized_params, case, d, V_lst):
    """Log-likelihoods of MANY parameter vectors in one batched device call: the objective of
    ``optimization_wrapper`` (optimizer.py:396-567: parameter vector -> derived times ->
    model build -> summed forward log-likelihood) for every row of ``arg_sets``
    (n_sets x len(optimized_params)), without the history / best-model side effects.
    One batched model build (FP64 tensor cores) and one multi-set forward sweep over the
    resident blocks; the building block of a batched-simplex or multi-start optimiser
    (SURVEY 8f N3).  Returns a float64 vector of length n_sets."""
    arg_sets = np.atleast_2d(np.asarray(arg_sets, dtype=np.float64))
    if arg_sets.shape[1] != len(optimized_params):
        raise ValueError("arg_sets must have one column per optimised parameter")
    rows = []
    for args in arg_sets:
        d_copy = d.copy()
        for i, param in enumerate(optimized_params):
            d_copy[param] = float(args[i])
        derive_times(d_copy, case)
        rows.append(model_args(d_copy))
    return _objective_batch(np.array(rows, dtype=np.float64), d["n_int_AB"], d["n_int_ABC"], V_lst)


def _optimizer_batched(optim_variables, optim_list, bounds, d, V_lst, res_name, case):
    """Nelder-Mead with one batched objective call per iteration (batched_simplex.py): the
    reflection / expansion / contraction candidates of an iteration are evaluated together by
    ``loglik_sweep``; the history and best-model files receive exactly the evaluations, in the
    order, the sequential search (optimizer.py:623-637) makes."""
    from .batched_simplex import minimize_neldermead_batched

    output_dir, output_prefix = os.path.split(res_name)
    history = os.path.join(output_dir, f"{output_prefix}.optimization_history.csv")
    best_model_yaml = os.path.join(output_dir, f"{output_prefix}.best_model.yaml")
    rank, _world = dist_.rank_world()
    info = {"Nfeval": 0, "time": time.time()}

    def consume(x, f):
        if rank == 0:
            write_list([info["Nfeval"]] + np.asarray(x).tolist() + [-f, time.time() - info["time"]], history)
            update_best_model(best_model_yaml, optim_variables, x, -f, info["Nfeval"])
        info["Nfeval"] += 1

    return minimize_neldermead_batched(
        lambda X: -loglik_sweep(X, optim_variables, case, d, V_lst),
        optim_list, bounds=bounds, maxiter=10000, consume=consume, disp=True)


def optimizer(optim_variables, optim_list, bounds, fixed_params, V_lst, res_name, case,
              method="Nelder-Mead", header=True):
    """scipy.optimize.minimize over the scaled parameters (optimizer.py:586-637)."""
    from scipy.optimize import minimize

    output_dir, output_prefix = os.path.split(res_name)
    history = os.path.join(output_dir, f"{output_prefix}.optimization_history.csv")
    rank, _world = dist_.rank_world()
    if header and rank == 0:
        write_list(["n_eval"] + list(optim_variables) + ["loglik", "time"], history)
    if method == "Nelder-Mead-batched":
        return _optimizer_batched(optim_variables, optim_list, bounds, fixed_params.copy(), V_lst, res_name, case)
    return minimize(
        optimization_wrapper,
        x0=optim_list,
        args=(optim_variables, case, fixed_params.copy(), V_lst, res_name,
              {"Nfeval": 0, "time": time.time()}),
        method=method,
        bounds=bounds,
        options={"maxiter": 10000, "disp": True},
    )
