"""Command-line workflows — the reference's three console scripts on the GPU path:

    itrails-optimize   workflow_optimize.py:19-493
    itrails-viterbi    workflow_viterbi.py:19-749
    itrails-posterior  workflow_posterior.py:19-721

Same arguments, same YAML layout (``fixed_parameters`` / ``optimized_parameters`` /
``settings``), same validation messages for the common mistakes, same unit scaling (times
and population sizes times ``mu``, ``r`` divided by ``mu``), same output files:
``<prefix>.starting_params.yaml``, ``.best_model.yaml``, ``.optimization_history.csv``,
``.hidden_states.csv`` (``_2`` if it exists), ``.viterbi.csv`` (run-length rows
``Block_idx,position_start,position_end,most_likely_state``) and ``.posterior.csv``
(``alignment_block_idx,position_idx,prob_state_0..``).  ``n_cpu`` is accepted and ignored
(one process drives one GPU).
"""
from __future__ import annotations

import argparse
import csv
import math
import os
import sys
from math import inf

import numpy as np
import yaml

from . import __version__
from .cutpoints import cutpoints_AB, cutpoints_ABC
from .ngpu import update_n_cpu, update_n_gpu
from .yaml_helpers import FlowSeq, load_config

TIME_CASES = {
    frozenset(["t_A", "t_B", "t_C"]), frozenset(["t_1", "t_A"]), frozenset(["t_1", "t_B"]),
    frozenset(["t_1", "t_C"]), frozenset(["t_A", "t_B"]), frozenset(["t_A", "t_C"]),
    frozenset(["t_B", "t_C"]), frozenset(["t_1"]),
}
TOPOLOGY_NAMES = {0: "({sp1,sp2},sp3)", 1: "((sp1,sp2),sp3)", 2: "((sp1,sp3),sp2)", 3: "((sp2,sp3),sp1)"}


def _pick_path(kind, cmd, cfg):
    """Command line wins over the config file (workflow_optimize.py:52-93)."""
    label = "MAF alignment file" if kind == "input" else "Output file"
    if cmd and cfg:
        print(f"Warning: {label} specified in both config file ({cfg}) and command-line ({cmd}). "
              f"Using command-line {kind}.")
        return cmd
    if cmd or cfg:
        return cmd or cfg
    raise ValueError(f"Error: {label} not specified in config file or command-line.")


def _positive_int(name, v):
    if not (isinstance(v, int) and not isinstance(v, bool) and v > 0):
        raise ValueError(f"{name} must be a positive integer")
    return v


def _scale(name, value, mu):
    return float(value) / mu if name == "r" else float(value) * mu


# ---------------------------------------------------------------------------
# itrails-optimize
# ---------------------------------------------------------------------------
def prepare_optimize(config, n_int_AB, n_int_ABC):
    """Validates the parameter section of an optimise config and returns
    ``(optim_variables, optim_list, bounds_list, fixed_dict, case)`` in the scaled units
    the objective works in (workflow_optimize.py:110-405)."""
    fixed_params = config.get("fixed_parameters") or {}
    optimized_params = config.get("optimized_parameters") or {}
    mu = float(fixed_params["mu"])
    if not mu > 0:
        raise ValueError("mu must be a positive float or int.")
    fixed_dict = {"n_int_AB": _positive_int("n_int_AB", n_int_AB),
                  "n_int_ABC": _positive_int("n_int_ABC", n_int_ABC)}
    optim_variables, optim_list, bounds_list = [], [], []

    def take(param, required):
        if param in fixed_params and param in optimized_params:
            raise ValueError(f"Parameter '{param}' cannot be both fixed and optimized.")
        if param in fixed_params:
            fixed_dict[param] = fixed_params[param]
            return True
        if param in optimized_params:
            start, lo, hi = optimized_params[param]
            optim_variables.append(param)
            optim_list.append(start)
            bounds_list.append((lo, hi))
            return True
        if required:
            raise ValueError("Parameters 't_2', 'N_ABC', 'N_AB' and 'r' must be present in optimized or fixed parameters.")
        return False

    found = {p for p in ("t_1", "t_A", "t_B", "t_C") if take(p, False)}
    if frozenset(found) not in TIME_CASES:
        raise ValueError(f"Invalid combination of time values: {found}, check possible combinations in the documentation.")
    case = frozenset(found)
    for p in ("t_2", "N_ABC", "N_AB", "r"):
        take(p, True)

    if "t_upper" in optimized_params or "t_upper" in fixed_params:
        take("t_upper", False)
        tu = [optim_list[-1], *bounds_list[-1]] if "t_upper" in optimized_params else [fixed_dict["t_upper"]]
        if any(float(x) < 0 for x in tu):
            raise ValueError("Parameter 't_upper' cannot be negative. Please check your input parameters.")
    else:
        # t_upper from t_3 and N_ABC (workflow_optimize.py:247-331)
        print("Warning: 't_upper' not found in parameter definition. Calculating from 't_3' and 'N_ABC'.")
        last = lambda N: cutpoints_ABC(n_int_ABC, 1 / N)[-2]
        if "N_ABC" in optimized_params:
            N0, Nlo, Nhi = optimized_params["N_ABC"]
        elif "N_ABC" in fixed_params:
            N0 = Nlo = Nhi = fixed_params["N_ABC"]
        else:
            raise ValueError("'N_ABC' not found in parameter definition.")
        if "t_3" in optimized_params:
            t0, tlo, thi = optimized_params["t_3"]
        elif "t_3" in fixed_params:
            if "N_ABC" in fixed_params:
                raise ValueError("At least one, 't_3' or 'N_ABC' must be present in optimized parameters.")
            t0 = tlo = thi = fixed_params["t_3"]
        else:
            raise ValueError("'t_3' not found in parameter definition.")
        start, lo, hi = t0 - last(N0), tlo - last(Nhi), thi - last(Nlo)
        if not (lo <= start <= hi):
            raise ValueError(f"When calculating t_upper from t_3 and N_ABC, the starting value ({start}) was not between "
                             f"the minimum ({lo}) and maximum ({hi}).")
        if min(start, lo, hi) < 0:
            raise ValueError("Calculated 't_upper' values cannot be negative. Please check your input parameters.")
        optim_variables.append("t_upper")
        optim_list.append(start)
        bounds_list.append((lo, hi))

    if "t_out" in fixed_params:
        fixed_dict["t_out"] = fixed_params["t_out"]
    elif "t_out" in optimized_params:
        raise ValueError("Parameter 't_out' has to be fixed.")

    for i, param in enumerate(optim_variables):
        start, lo, hi = float(optim_list[i]), float(bounds_list[i][0]), float(bounds_list[i][1])
        if not (lo <= start <= hi):
            raise ValueError(f"Starting value for '{param}' ({start}) must be between the minimum ({lo}) and maximum ({hi}).")
        if start <= 0:
            raise ValueError(f"Starting value for '{param}' must be a positive number.")
        if lo <= 0:
            raise ValueError(f"Minimum value for '{param}' must be a positive number.")
        optim_list[i] = _scale(param, start, mu)
        bounds_list[i] = (_scale(param, lo, mu), _scale(param, hi, mu))
    for param in list(fixed_dict):
        if param not in ("n_int_AB", "n_int_ABC"):
            fixed_dict[param] = _scale(param, fixed_dict[param], mu)
    return optim_variables, optim_list, bounds_list, fixed_dict, case


def optimize_main(argv=None):
    from .optimizer import optimizer
    from .read_data import maf_parser

    parser = argparse.ArgumentParser(
        description="Optimize workflow using TRAILS (B200 path)",
        usage="itrails-optimize <config.yaml> --input PATH_MAF --output OUTPUT_PATH")
    parser.add_argument("--version", action="version", version=f"%(prog)s {__version__}")
    parser.add_argument("config_file", type=str, help="Path to the YAML config file.")
    parser.add_argument("--input", type=str, help="Path to the MAF alignment file.")
    parser.add_argument("--output", type=str, help="Path and prefix for output files: 'directory/prefix'.")
    args = parser.parse_args(argv)

    config = load_config(args.config_file)
    settings = config["settings"]
    maf_path = _pick_path("input", args.input, settings.get("input_maf"))
    if not (args.input and settings.get("input_maf")):
        print(f"Using MAF alignment file: {maf_path}")
    user_output = _pick_path("output", args.output, settings.get("output_prefix"))
    output_dir, output_prefix = os.path.split(user_output)
    os.makedirs(output_dir or ".", exist_ok=True)
    print(f"Results will be saved to: {output_dir}.")
    settings["n_cpu"] = update_n_cpu(settings.get("n_cpu"))
    if settings.get("n_gpu") is not None:        # (this package's addition: GPUs driven by this process)
        update_n_gpu(settings.get("n_gpu"))
    settings["output_prefix"], settings["input_maf"] = user_output, maf_path
    mu = float(config["fixed_parameters"]["mu"])
    method = str(settings["method"]).lower()
    # 'nelder-mead-batched' is this package's addition: the same simplex search with the
    # candidates of an iteration evaluated in one batched device call (batched_simplex.py)
    methods = {"nelder-mead": "Nelder-Mead", "l-bfgs-b": "L-BFGS-B", "nelder-mead-batched": "Nelder-Mead-batched"}
    if method not in methods:
        raise ValueError("Method must be one of ['nelder-mead', 'l-bfgs-b'].")
    print(f"Using optimization method: {method}")

    optim_variables, optim_list, bounds_list, fixed_dict, case = prepare_optimize(
        config, settings["n_int_AB"], settings["n_int_ABC"])

    unscale = lambda p, v: float(v) * mu if p == "r" else float(v) / mu
    fixed_user = {k: unscale(k, v) for k, v in fixed_dict.items() if k not in ("n_int_AB", "n_int_ABC")}
    fixed_user["mu"] = mu
    if "species_list" in settings:
        settings["species_list"] = FlowSeq(settings["species_list"])
    starting = {
        "fixed_parameters": fixed_user,
        "optimized_parameters": {p: FlowSeq([unscale(p, v), unscale(p, b[0]), unscale(p, b[1])])
                                 for p, v, b in zip(optim_variables, optim_list, bounds_list)},
        "settings": settings,
    }
    with open(os.path.join(output_dir, f"{output_prefix}.starting_params.yaml"), "w") as fh:
        yaml.dump(starting, fh, default_flow_style=False)
    best_model_yaml = os.path.join(output_dir, f"{output_prefix}.best_model.yaml")
    with open(best_model_yaml, "w") as fh:
        yaml.dump({"fixed_parameters": fixed_user, "optimized_parameters": {},
                   "results": {"log_likelihood": -inf, "iteration": None}, "settings": settings}, fh)

    maf_alignment = maf_parser(maf_path, list(settings["species_list"]))
    if not maf_alignment:
        raise ValueError("Error reading MAF alignment file.")
    print("Running optimization...")
    res = optimizer(optim_variables=optim_variables, optim_list=optim_list, bounds=bounds_list,
                    fixed_params=fixed_dict, V_lst=maf_alignment, res_name=user_output, case=case,
                    method=methods[method], header=True)
    print(f"Optimization complete. Results saved to "
          f"{os.path.join(output_dir, f'{output_prefix}.optimization_history.csv')}.\n"
          f" Best model saved to {best_model_yaml}.")
    return res


# ---------------------------------------------------------------------------
# itrails-viterbi / itrails-posterior
# ---------------------------------------------------------------------------
def _decode_parser(what):
    p = argparse.ArgumentParser(
        description=f"Run {what} decoding using iTRAILS (B200 path)",
        usage=f"itrails-{what} --config-file CONFIG_FILE --input PATH_MAF --output OUTPUT_PATH --PARAMETERS")
    p.add_argument("--version", action="version", version=f"%(prog)s {__version__}")
    p.add_argument("--config-file", type=str)
    p.add_argument("--input", type=str)
    p.add_argument("--output", type=str)
    for name in ("mu", "t1", "t_A", "t_B", "t_C", "t2", "t3", "t_upper", "t_out", "N_AB", "N_ABC", "r"):
        p.add_argument(f"--{name}", type=float)
    p.add_argument("--n_cpu", type=int)
    p.add_argument("--n_gpu", type=int, help="GPUs this process drives (blocks are sharded over them); default 1")
    p.add_argument("--species_list", nargs="+")
    p.add_argument("--reference", type=str)
    p.add_argument("--n_int_AB", type=int)
    p.add_argument("--n_int_ABC", type=int)
    p.add_argument("--cutpoints_AB", nargs="+", type=float)
    p.add_argument("--cutpoints_ABC", nargs="+", type=float)
    return p


def prepare_decode(config):
    """Fixed-parameter model description of the decoding workflows
    (workflow_viterbi.py:208-597): returns ``(fixed_dict, norm_cut_AB, norm_cut_ABC,
    abs_cut_AB, abs_cut_ABC)``; ``fixed_dict`` holds the nine scaled scalars of
    ``trans_emiss_calc`` plus the discretisation."""
    settings = config["settings"]
    cut_AB, cut_ABC = settings.get("cutpoints_AB"), settings.get("cutpoints_ABC")
    n_int_AB, n_int_ABC = settings.get("n_int_AB"), settings.get("n_int_ABC")
    if not n_int_AB and not cut_AB:
        raise ValueError("Error: n_int_AB must be specified in the config file for automatic cutpoints, n_int_AB and "
                         "cutpoints_AB must be specified in the config file for manual cutpoints.")
    if not n_int_ABC and not cut_ABC:
        raise ValueError("Error: n_int_ABC must be specified in the config file for automatic cutpoints, n_int_ABC and "
                         "cutpoints_ABC must be specified in the config file for manual cutpoints.")
    if cut_AB and n_int_AB and len(cut_AB) != n_int_AB + 1:
        raise ValueError("Error: cutpoints_AB must have n_int_AB + 1 values, check the config file.")
    if cut_ABC and n_int_ABC and len(cut_ABC) != n_int_ABC:
        raise ValueError("Error: cutpoints_ABC must have n_int_ABC values, check the config file.")
    n_int_AB = _positive_int("n_int_AB", n_int_AB if n_int_AB else len(cut_AB) - 1)
    n_int_ABC = _positive_int("n_int_ABC", n_int_ABC if n_int_ABC else len(cut_ABC))
    fixed_params = dict(config.get("fixed_parameters") or {})
    optimized_params = dict(config.get("optimized_parameters") or {})
    mu = float(fixed_params["mu"])
    if not mu > 0:
        raise ValueError("mu must be a positive float or int.")
    for p in optimized_params:
        if p in fixed_params:
            raise ValueError(f"Parameter '{p}' cannot be both fixed and optimized.")
    if "t_out" in optimized_params:
        raise ValueError("Parameter 't_out' has to be fixed.")
    vals = {**optimized_params, **{k: v for k, v in fixed_params.items() if k != "mu"}}
    for p in ("t_2", "N_ABC", "N_AB", "r"):
        if p not in vals:
            raise ValueError("Parameters 't_2', 'N_ABC', 'N_AB' and 'r' must be present in optimized or fixed parameters.")
    found = {p for p in ("t_1", "t_A", "t_B", "t_C") if p in vals}
    if frozenset(found) not in TIME_CASES:
        raise ValueError(f"Invalid combination of time values: {found}, check possible combinations in the documentation.")
    case = frozenset(found)
    if "t_A" in vals:
        pre_t_A = float(vals["t_A"])
    elif "t_1" in vals:
        pre_t_A = float(vals["t_1"])
    else:
        raise ValueError("t_A or t_1 is needed to place the cutpoints.")
    pre_t_2, pre_N_AB, pre_N_ABC = float(vals["t_2"]), float(vals["N_AB"]), float(vals["N_ABC"])

    if cut_AB is None:
        abs_cut_AB = [float(x) for x in pre_t_A + cutpoints_AB(n_int_AB, pre_t_2, 1 / pre_N_AB)]
    else:
        abs_cut_AB = [float(x) for x in cut_AB]
    norm_cut_AB = [(x - pre_t_A) / pre_N_ABC for x in abs_cut_AB]
    if cut_ABC is None:
        norm_cut_ABC = [float(x) for x in cutpoints_ABC(n_int_ABC, 1)]
        abs_cut_ABC = [x * pre_N_ABC + pre_t_A + pre_t_2 for x in norm_cut_ABC]
    else:
        abs_cut_ABC = [float(x) for x in cut_ABC]
        norm_cut_ABC = [(x - pre_t_A - pre_t_2) / pre_N_ABC for x in abs_cut_ABC] + [float("inf")]
        abs_cut_ABC = abs_cut_ABC + [float("inf")]

    if "t_upper" not in vals:
        print("Warning: 't_upper' not found in parameter definition. Calculating from 't_3' and 'N_ABC'.")
        if "t_3" not in vals:
            raise ValueError("'t_3' not found in parameter definition.")
        vals["t_upper"] = float(vals["t_3"]) - norm_cut_ABC[-2] * pre_N_ABC
    vals.pop("t_3", None)
    fixed_dict = {"n_int_AB": n_int_AB, "n_int_ABC": n_int_ABC}
    for p, v in vals.items():
        if p != "t_upper" and not float(v) > 0:
            raise ValueError(f"Value for '{p}' must be a positive number.")
        fixed_dict[p] = _scale(p, v, mu)
    if fixed_dict["t_upper"] < 0:
        raise ValueError("Parameter 't_upper' must be a positive number. "
                         f"Given/calculated value: {fixed_dict['t_upper']}")
    # derived times, t_out (workflow_viterbi.py:429-560 == optimizer.py:419-541)
    d = fixed_dict
    tail = norm_cut_ABC[-2] * d["N_ABC"] + d["t_upper"] + 2 * d["N_ABC"]
    if "t_1" in case:
        t_1 = d.pop("t_1")
        d.setdefault("t_A", t_1)
        d.setdefault("t_B", t_1)
        d.setdefault("t_C", t_1 + d["t_2"])
        t_out = t_1 + d["t_2"] + tail
    else:
        if case == frozenset(["t_A", "t_B"]):
            d["t_C"] = (d["t_A"] + d["t_B"]) / 2 + d["t_2"]
        elif case == frozenset(["t_A", "t_C"]):
            d["t_B"] = (d["t_A"] + d["t_C"] - d["t_2"]) / 2
        elif case == frozenset(["t_B", "t_C"]):
            d["t_A"] = (d["t_B"] + d["t_C"] - d["t_2"]) / 2
        t_out = (((d["t_A"] + d["t_B"]) / 2 + d["t_2"]) + d["t_C"]) / 2 + tail
    d.setdefault("t_out", t_out)

    close = lambda x, y: math.isclose(x, y, rel_tol=1e-9, abs_tol=1e-12)
    lo, hi = pre_t_A, pre_t_A + pre_t_2
    if (abs_cut_AB[0] < lo and not close(abs_cut_AB[0], lo)) or (abs_cut_AB[-1] > hi and not close(abs_cut_AB[-1], hi)):
        raise ValueError(f"cutpoints_AB must lie within [t_A, t_A + t_2].Given cutpoints_AB: {abs_cut_AB}, "
                         f"t_A: {lo}, t_A + t_2: {hi}.")
    lo, hi = pre_t_A + pre_t_2, d["t_out"] / mu
    if (abs_cut_ABC[0] < lo and not close(abs_cut_ABC[0], lo)) or (abs_cut_ABC[-2] > hi and not close(abs_cut_ABC[-2], hi)):
        raise ValueError(f"cutpoints_ABC must lie within [t_A + t_2, t_out].Given cutpoints_ABC: {abs_cut_ABC}, "
                         f"t_A + t_2: {lo}, t_out: {hi}.")
    return fixed_dict, norm_cut_AB, norm_cut_ABC, abs_cut_AB, abs_cut_ABC


def write_hidden_states(path, hidden_names, abs_cut_ABC):
    """workflow_viterbi.py:636-684 (intervals are printed from the ABC cutpoints for both
    coalescences, as the reference does)."""
    with open(path, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["state_idx", "topology", "interval_1st_coalescent", "interval_2nd_coalescent", "shorthand_name"])
        for idx, sh in hidden_names.items():
            i1 = f"{abs_cut_ABC[sh[1]]:.2f}-{abs_cut_ABC[sh[1] + 1]:.2f}"
            i2 = f"{abs_cut_ABC[sh[2]]:.2f}-{abs_cut_ABC[sh[2] + 1]:.2f}"
            w.writerow([idx, TOPOLOGY_NAMES.get(sh[0], "Unknown"), i1, i2, sh])


def viterbi_segments(res, coords=None):
    """Run-length rows ``(start, end, state)`` of one block's path
    (workflow_viterbi.py:698-743); with ``coords`` the positions are reference
    coordinates and -9 columns are skipped over."""
    res = np.asarray(res)
    n = len(res)
    if n == 0:
        return []
    if coords is None:
        cuts = np.flatnonzero(res[1:] != res[:-1]) + 1
        starts = np.concatenate(([0], cuts))
        ends = np.concatenate((cuts - 1, [n - 1]))
        return [(int(s), int(e), res[s]) for s, e in zip(starts, ends)]
    first = next((i for i, x in enumerate(coords) if x != -9), None)
    if first is None:
        return []
    rows = []
    seg_start = cur_nn = coords[first]
    cur_state = res[first]
    for pos in range(first, n):
        if seg_start == -9:
            seg_start = coords[pos]
            cur_state = res[pos]
            cur_nn = seg_start
            continue
        if res[pos] != cur_state:
            rows.append((seg_start, cur_nn, cur_state))
            seg_start = coords[pos]
            cur_state = res[pos]
        cur_nn = coords[pos] if coords[pos] != -9 else cur_nn
    if not (seg_start == cur_nn == -9):
        rows.append((seg_start, cur_nn, cur_state))
    return rows


def _decode_main(what, argv):
    from .get_trans_emiss import trans_emiss_calc
    from .optimizer import post_prob_to_csv, post_prob_wrapper, viterbi_wrapper
    from .read_data import maf_parser, parse_coordinates

    parser = _decode_parser(what)
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) == 0:
        parser.print_usage()
        sys.exit("Error: No arguments provided. Please provide either a config file, command-line parameters, or both.")
    args = parser.parse_args(argv)
    config = {"fixed_parameters": {}, "optimized_parameters": {}, "settings": {}}
    if args.config_file:
        config = load_config(args.config_file)
        for k in ("fixed_parameters", "optimized_parameters", "settings"):
            config[k] = config.get(k) or {}
    if args.mu is not None:
        config["fixed_parameters"]["mu"] = args.mu
    elif "mu" not in config["fixed_parameters"]:
        raise ValueError("Error: mu must be specified either in config file or via --mu")
    cli = {"t_1": args.t1, "t_A": args.t_A, "t_B": args.t_B, "t_C": args.t_C, "t_2": args.t2, "t_3": args.t3,
           "t_upper": args.t_upper, "t_out": args.t_out, "N_AB": args.N_AB, "N_ABC": args.N_ABC, "r": args.r}
    for p, v in cli.items():
        if v is not None:
            config["optimized_parameters"].pop(p, None)
            config["fixed_parameters"][p] = v
    for key in ("n_cpu", "n_gpu", "species_list", "reference", "n_int_AB", "n_int_ABC", "cutpoints_AB", "cutpoints_ABC"):
        if getattr(args, key) is not None:
            config["settings"][key] = getattr(args, key)
    settings = config["settings"]
    maf_path = _pick_path("input", args.input, settings.get("input_maf"))
    user_output = _pick_path("output", args.output, settings.get("output_prefix"))
    output_dir, output_prefix = os.path.split(user_output)
    os.makedirs(output_dir or ".", exist_ok=True)
    print(f"Results will be saved to: {output_dir} as '{output_prefix}.{what}.csv'.")
    update_n_cpu(settings.get("n_cpu"))
    if settings.get("n_gpu") is not None:
        update_n_gpu(settings.get("n_gpu"))
    if "species_list" not in settings or len(settings["species_list"]) != 4:
        raise ValueError("Error: species_list must name four species.")
    species_list = list(settings["species_list"])
    mu = float(config["fixed_parameters"]["mu"])

    fixed_dict, norm_cut_AB, norm_cut_ABC, abs_cut_AB, abs_cut_ABC = prepare_decode(config)
    print("Parameters validated:")
    print(f"Cutpoints AB: {abs_cut_AB}")
    print(f"Cutpoints ABC: {abs_cut_ABC}")
    for key, value in fixed_dict.items():
        if key in ("n_int_AB", "n_int_ABC"):
            print(f"{key}: {value}")
        else:
            print(f"{key}: {value * mu if key == 'r' else value / mu}")

    print("Reading MAF alignment file.")
    maf_alignment = maf_parser(maf_path, species_list)
    if not maf_alignment:
        raise ValueError("Error reading MAF alignment file.")
    reference = settings.get("reference")
    ref_coordinates = parse_coordinates(maf_path, species_list, reference) if reference is not None else None

    print("Calculating transition and emission probability matrices.")
    a, b, pi, hidden_names, _observed = trans_emiss_calc(
        fixed_dict["t_A"], fixed_dict["t_B"], fixed_dict["t_C"], fixed_dict["t_2"], fixed_dict["t_upper"],
        fixed_dict["t_out"], fixed_dict["N_AB"], fixed_dict["N_ABC"], fixed_dict["r"],
        fixed_dict["n_int_AB"], fixed_dict["n_int_ABC"], norm_cut_AB, norm_cut_ABC)

    # Several GPUs: every process decodes its share of the blocks, the wrappers return the
    # whole result in input order, and rank 0 alone writes the files.
    from . import distributed as dist_
    writer = dist_.rank_world()[0] == 0
    hidden_file = os.path.join(output_dir, f"{output_prefix}.hidden_states.csv")
    if writer:
        if os.path.exists(hidden_file):
            print(f"Warning: File '{hidden_file}' already exists.")
            hidden_file = os.path.join(output_dir, f"{output_prefix}.hidden_states_2.csv")
            print(f"Using an alternative file name: {hidden_file}")
        write_hidden_states(hidden_file, hidden_names, abs_cut_ABC)
        print(f"Hidden states written to file {hidden_file}.")

    output_file = os.path.join(output_dir, f"{output_prefix}.{what}.csv")
    if what == "viterbi":
        print("Running viterbi.")
        result = viterbi_wrapper(a=a, b=b, pi=pi, V_lst=maf_alignment)
        if not writer:
            return output_file
        print("Writing results to file.")
        with open(output_file, "w", newline="") as fh:
            w = csv.writer(fh)
            w.writerow(["Block_idx", "position_start", "position_end", "most_likely_state"])
            for block_idx, res in enumerate(result):
                coords = ref_coordinates[block_idx] if ref_coordinates is not None else None
                for s, e, state in viterbi_segments(res, coords):
                    w.writerow([block_idx, s, e, state])
        print(f"Viterbi decoding complete. Results saved to {output_file}.")
    else:
        print("Running posterior decoding.")
        native = os.environ.get("ITRAILS_PY_CSV") is None and (
            ref_coordinates is None or [len(c) for c in ref_coordinates] == [len(v) for v in maf_alignment])
        if native:
            print("Writing results to file.")
            post_prob_to_csv(a, b, pi, maf_alignment, output_file, ref_coordinates, settings.get("n_cpu") or 0)
            print(f"Posterior decoding complete. Results saved to {output_file}.")
            return output_file
        result = post_prob_wrapper(a=a, b=b, pi=pi, V_lst=maf_alignment)
        if not writer:
            return output_file
        print("Writing results to file.")
        with open(output_file, "w", newline="") as fh:
            w = csv.writer(fh)
            n_states = result[0].shape[1] if result else 0
            w.writerow(["alignment_block_idx", "position_idx"] + [f"prob_state_{i}" for i in range(n_states)])
            for block_idx, arr in enumerate(result):
                pos = ref_coordinates[block_idx] if ref_coordinates is not None else range(len(arr))
                for p, row in zip(pos, arr):
                    w.writerow([block_idx, p] + row.tolist())
        print(f"Posterior decoding complete. Results saved to {output_file}.")
    return output_file


def viterbi_main(argv=None):
    return _decode_main("viterbi", argv)


def posterior_main(argv=None):
    return _decode_main("posterior", argv)
