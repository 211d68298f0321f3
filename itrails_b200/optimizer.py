"""HMM wrappers and optimiser driver — same names, arguments and return types as the
reference's optimizer.py, with the numerics on the GPU:

    loglik_wrapper / loglik_wrapper_par   optimizer.py:40-116
    post_prob_wrapper                     optimizer.py:241-262
    viterbi_wrapper                       optimizer.py:357-377
    optimization_wrapper / optimizer      optimizer.py:396-637
    write_list                            optimizer.py:380-393

Under ``torch.distributed`` (one process per GPU) the log-likelihood wrappers shard
the blocks over ranks and all-reduce the scalar; the decoders return this rank's
blocks only (see ``itrails_b200.distributed``).
"""
from __future__ import annotations

import os
import time

import numpy as np

from . import distributed as dist_
from .cutpoints import cutpoints_ABC
from .engine_cache import ensure_blocks, get_engine
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
        first = np.array([int(V[0]) for V in V_lst])
        omega0 = np.log(pi[None, :] * E[:, first].T)
    return log_a, log_E, np.ascontiguousarray(omega0)


# ---------------------------------------------------------------------------
# wrappers
# ---------------------------------------------------------------------------
def _local(V_lst):
    if dist_.is_active():
        local, ids = dist_.shard_blocks(V_lst)
        return local, ids
    return V_lst, None


_SHARD_CACHE = {}


def _resident(V_lst):
    """Engine with this rank's share of V_lst resident in HBM."""
    if not dist_.is_active():
        return ensure_blocks(V_lst), V_lst
    key = id(V_lst)
    hit = _SHARD_CACHE.get(key)
    if hit is None or hit[0] is not V_lst:
        _SHARD_CACHE.clear()
        local, ids = dist_.shard_blocks(V_lst)
        _SHARD_CACHE[key] = (V_lst, local, ids)
        hit = _SHARD_CACHE[key]
    return ensure_blocks(hit[1]), hit[1]


def loglik_wrapper(a, b, pi, V_lst):
    """Sum over blocks of the forward log-likelihood (optimizer.py:93-116)."""
    eng, _ = _resident(V_lst)
    eng.set_model(a, b, pi)
    total = eng.loglik()
    if dist_.is_active():
        total = dist_.allreduce_sum(total, eng.device)
    return float(total[0])      # a Python float, as the reference's numba forward_loglik returns


def loglik_wrapper_par(a, b, pi, V_lst):
    """optimizer.py:40-65 — the reference farms blocks to joblib workers; here every
    block is a warp-level chain on the GPU, so this is ``loglik_wrapper``."""
    return loglik_wrapper(a, b, pi, V_lst)


def post_prob_wrapper(a, b, pi, V_lst):
    """List of (T, K) float64 posterior matrices, one per block (optimizer.py:241-262)."""
    eng, local = _resident(V_lst)
    eng.set_model(a, b, pi)
    post = eng.posterior()
    return eng.split(post)


def post_prob_to_csv(a, b, pi, V_lst, output_file, ref_coordinates=None, n_threads=0):
    """Posterior decoding written straight to ``output_file`` in the reference's format
    (workflow_posterior.py:697-716) without materialising the list of (T, K) matrices on
    the host: the result stays in HBM and the native writer streams it block by block.
    Byte-identical to ``csv.writer`` over ``post_prob_wrapper``'s result."""
    eng, local = _resident(V_lst)
    eng.set_model(a, b, pi)
    positions = None
    if ref_coordinates is not None:
        if [len(c) for c in ref_coordinates] != [len(v) for v in local]:
            raise IndexError("reference coordinates do not match the alignment blocks")
        positions = np.concatenate([np.asarray(c, dtype=np.int64) for c in ref_coordinates])
    eng.posterior(fetch=False)
    eng.write_posterior_csv(output_file, positions, n_threads)


def viterbi_wrapper(a, b, pi, V_lst):
    """List of float64 state paths, one per block (optimizer.py:357-377; the reference
    returns float arrays, e.g. ``14.0``)."""
    eng, local = _resident(V_lst)
    eng.set_model(a, b, pi)
    log_a, log_E, omega0 = viterbi_tables(a, b, pi, local)
    path = eng.viterbi(log_a, log_E, omega0)
    return [p.astype(np.float64) for p in eng.split(path)]


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
    eng, _ = _resident(V_lst)
    eng.build_model(np.array([model_args(d_copy)]), d_copy["n_int_AB"], d_copy["n_int_ABC"], fetch=False)
    total = eng.loglik()
    if dist_.is_active():
        total = dist_.allreduce_sum(total, eng.device)
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
    eng, _ = _resident(V_lst)
    eng.build_model(np.array(rows, dtype=np.float64), d["n_int_AB"], d["n_int_ABC"], fetch=False)
    total = eng.loglik()
    if dist_.is_active():
        total = dist_.allreduce_sum(total, eng.device)
    return np.asarray(total, dtype=np.float64)


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
