"""YAML config I/O with the reference's file semantics (yaml_helpers.py:7-118):
``load_config`` exits on unreadable files, ``FlowSeq`` lists dump inline, and
``update_best_model`` rewrites ``<prefix>.best_model.yaml`` when the log-likelihood
improves, converting parameters back to user units (divide by mu; ``r`` times mu)."""
import os
import sys

import yaml


class FlowSeq(list):
    """A list that PyYAML writes in flow style (``[a, b, c]``)."""


yaml.add_representer(
    FlowSeq, lambda dumper, data: dumper.represent_sequence("tag:yaml.org,2002:seq", data, flow_style=True))


def load_config(config_file):
    try:
        with open(config_file, "r") as fh:
            return yaml.safe_load(fh)
    except Exception as exc:  # same behaviour as the reference: report and exit(1)
        print(f"Error loading config file: {exc}", file=sys.stderr)
        sys.exit(1)


def update_best_model(best_model_yaml, optim_variables, current_optim_params, current_result, iteration):
    if not os.path.exists(best_model_yaml):
        raise FileNotFoundError(f"Best model file not found: {best_model_yaml}")
    with open(best_model_yaml, "r") as fh:
        try:
            data = yaml.safe_load(fh)
        except yaml.YAMLError as exc:
            print(f"Error loading best model file: {exc}")
            sys.exit(1)
    mu = float(data["fixed_parameters"]["mu"])
    best = data["results"]["log_likelihood"]
    if best is not None and not current_result > best:
        return
    params = {}
    for name, value in zip(optim_variables, current_optim_params):
        params[name] = float(value) * mu if name == "r" else float(value) / mu
    data["optimized_parameters"] = params
    data["results"]["log_likelihood"] = float(current_result)
    data["results"]["iteration"] = iteration
    with open(best_model_yaml, "w") as fh:
        yaml.dump(data, fh)
