#!/usr/bin/env python
"""Generate golden vectors by RUNNING THE REFERENCE (trails-phylogeny/itrails).

This script is the only place the repo touches /root/reference.  It cannot run on
the GPU box (the reference is not there); its outputs are committed as small
``.npz`` fixtures in this directory and are what ``tests/`` compares against.

Run (CPU container only)::

    PYTHONPATH=/root/reference/src:oracle/_stubs NUMBA_CACHE_DIR=/tmp/nbcache \
        python tests/golden/make_golden.py symbols statespace model:1:1:example ...

Jobs
----
symbols                    625 observed strings + the `order` index lists
                           (reference: read_data.py:6-67)
statespace                 transitions / omega masks / state dicts for 1,2,3 species
                           (reference: trans_mat.py:577-598)
model:<nAB>:<nABC>:<tag>   (a, b, pi, hidden states) from trans_emiss_calc
                           (reference: get_trans_emiss.py:8-170)
recursions:<model file>    forward loglik / alpha / beta / posterior / Viterbi on
                           seeded blocks (reference: optimizer.py:146-377)
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# ---------------------------------------------------------------------------
# Parameter sets (natural units, as a user writes them in the YAML).
# "example" is examples/example_config.yaml's starting values (case {t_1}).
# The others exercise t_A != t_B != t_C (case {t_A,t_B,t_C}) and N_AB != N_ABC.
# ---------------------------------------------------------------------------
PARAM_SETS = {
    "example": dict(mu=1e-8, N_AB=50000.0, N_ABC=50000.0, t_1=240000.0, t_2=40000.0,
                    t_upper=745069.3855, r=1e-8),
    "asym1": dict(mu=1e-8, N_AB=30000.0, N_ABC=70000.0, t_A=200000.0, t_B=260000.0,
                  t_C=310000.0, t_2=55000.0, t_upper=600000.0, r=2e-8),
    "asym2": dict(mu=2e-8, N_AB=120000.0, N_ABC=20000.0, t_A=150000.0, t_B=90000.0,
                  t_C=200000.0, t_2=20000.0, t_upper=300000.0, r=5e-9),
    "highrec": dict(mu=1e-8, N_AB=50000.0, N_ABC=100000.0, t_1=100000.0, t_2=150000.0,
                    t_upper=1500000.0, r=8e-8),
    # t_2 / N_AB = 80: the AB interval is 80 coalescent units long, so the last AB cutpoint
    # (truncexpon.ppf at q = 1) is the upper end of the support — the closed form
    # -log1p(q expm1(-b)) / c overflows there (cutpoints.py:22-25)
    "longab": dict(mu=1e-8, N_AB=5000.0, N_ABC=60000.0, t_1=200000.0, t_2=400000.0,
                   t_upper=700000.0, r=1.5e-8),
}


def scaled_args(p, n_int_ABC):
    """Natural-unit parameters -> the 9 scaled scalars trans_emiss_calc takes.

    Follows workflow_optimize.py:369-405 (times, N multiplied by mu; r divided by
    mu) and optimizer.py:419-541 (t_out for case {t_1} and case {t_A,t_B,t_C})."""
    mu = p["mu"]
    N_AB, N_ABC = p["N_AB"] * mu, p["N_ABC"] * mu
    t_2, t_upper, r = p["t_2"] * mu, p["t_upper"] * mu, p["r"] / mu
    cut_last = -np.log(1.0 - (n_int_ABC - 1) / n_int_ABC)  # expon.ppf((n-1)/n), rate 1
    if "t_1" in p:
        t_A = t_B = p["t_1"] * mu
        t_C = t_A + t_2
        t_out = t_A + t_2 + cut_last * N_ABC + t_upper + 2 * N_ABC
    else:
        t_A, t_B, t_C = p["t_A"] * mu, p["t_B"] * mu, p["t_C"] * mu
        t_out = (((t_A + t_B) / 2 + t_2) + t_C) / 2 + cut_last * N_ABC + t_upper + 2 * N_ABC
    return dict(t_A=t_A, t_B=t_B, t_C=t_C, t_2=t_2, t_upper=t_upper, t_out=t_out,
                N_AB=N_AB, N_ABC=N_ABC, r=r)


def _ref():
    import itrails.ncpu as ncpu
    ncpu.update_n_cpu(1)


def job_symbols():
    _ref()
    from itrails.read_data import get_idx_state, get_obs_state_dct
    names = list(get_obs_state_dct())
    order = [np.asarray(get_idx_state(i), dtype=np.int64) for i in range(625)]
    off = np.zeros(626, dtype=np.int64)
    off[1:] = np.cumsum([len(o) for o in order])
    np.savez_compressed(os.path.join(HERE, "symbols.npz"),
                        names=np.array(names), order_offsets=off,
                        order_values=np.concatenate(order))


def job_statespace():
    _ref()
    from itrails.trans_mat import wrapper_state_general
    out = {}
    for n in (1, 2, 3):
        tr, omega, sd, nonrev = wrapper_state_general(n)
        out[f"transitions_{n}"] = np.asarray(tr, dtype=np.int64)
        keys = sorted(omega.keys())
        out[f"omega_keys_{n}"] = np.array(keys, dtype=np.int64)
        out[f"omega_masks_{n}"] = np.array([np.asarray(omega[k]) for k in keys])
        st = sorted(sd.items(), key=lambda kv: kv[1])
        out[f"states_{n}"] = np.array([k for k, _ in st], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "statespace.npz"), **out)


def job_model(n_ab, n_abc, tag):
    _ref()
    from itrails.get_trans_emiss import trans_emiss_calc
    s = scaled_args(PARAM_SETS[tag], n_abc)
    t0 = time.time()
    a, b, pi, hid, obs = trans_emiss_calc(
        s["t_A"], s["t_B"], s["t_C"], s["t_2"], s["t_upper"], s["t_out"],
        s["N_AB"], s["N_ABC"], s["r"], n_ab, n_abc, "standard", "standard")
    dt = time.time() - t0
    hidden = np.array([hid[i] for i in range(len(hid))], dtype=np.int64)
    observed = np.array([obs[i] for i in range(len(obs))])
    args = np.array([s[k] for k in ("t_A", "t_B", "t_C", "t_2", "t_upper", "t_out",
                                    "N_AB", "N_ABC", "r")])
    fn = os.path.join(HERE, f"model_{n_ab}_{n_abc}_{tag}.npz")
    np.savez_compressed(fn, args=args, n_int=np.array([n_ab, n_abc]), a=a, b=b, pi=pi,
                        hidden=hidden, observed=observed, ref_seconds=np.array(dt))
    print(f"{fn}: K={a.shape[0]} ref {dt:.1f}s", flush=True)


def sample_blocks(a, b, pi, lengths, seed, p_n=0.02):
    """Columns sampled from the HMM itself, then species randomly masked to N."""
    from itrails.read_data import get_obs_state_dct
    names = list(get_obs_state_dct())
    lookup = {s: i for i, s in enumerate(names)}
    rng = np.random.default_rng(seed)
    K = a.shape[0]
    blocks = []
    for T in lengths:
        z = rng.choice(K, p=pi / pi.sum())
        V = np.empty(T, dtype=np.int64)
        for t in range(T):
            if t:
                z = rng.choice(K, p=a[z] / a[z].sum())
            V[t] = rng.choice(256, p=b[z] / b[z].sum())
        for t in np.nonzero(rng.random(T) < p_n)[0]:
            s = list(names[V[t]])
            for k in rng.choice(4, size=rng.integers(1, 5), replace=False):
                s[k] = "N"
            V[t] = lookup["".join(s)]
        blocks.append(V)
    return blocks


def job_recursions(model_file, lengths=(1500, 1, 2, 3, 257, 700)):
    _ref()
    from numba.typed import List
    from itrails.optimizer import (backtrack_viterbi, backward, forward,
                                   forward_loglik, loglik_wrapper, post_prob,
                                   post_prob_wrapper, viterbi, viterbi_wrapper)
    from itrails.read_data import get_idx_state
    m = np.load(os.path.join(HERE, model_file))
    a, b, pi = m["a"], m["b"], m["pi"]
    blocks = sample_blocks(a, b, pi, lengths, seed=20261018)
    # a block that is all-N in places and one that starts with NNNN
    blocks[4][:3] = [624, 256, 300]
    order = List()
    for i in range(625):
        order.append(get_idx_state(i))
    out = {"model_file": np.array(model_file), "n_blocks": np.array(len(blocks))}
    tot = 0.0
    for i, V in enumerate(blocks):
        out[f"V_{i}"] = V
        ll = forward_loglik(a, b, pi, V, order)
        tot += ll
        out[f"loglik_{i}"] = np.array(ll)
        out[f"alpha_{i}"] = forward(a, b, pi, V, order)
        out[f"beta_{i}"] = backward(a, b, V, order)
        out[f"post_{i}"] = post_prob(a, b, pi, V, order)
        om, prev = viterbi(a, b, pi, V, order)
        out[f"vit_{i}"] = backtrack_viterbi(om, prev)
        out[f"vit_omega_last_{i}"] = om[-1]
    out["loglik_total"] = np.array(loglik_wrapper(a, b, pi, blocks))
    assert abs(out["loglik_total"] - tot) < 1e-6
    pw = post_prob_wrapper(a, b, pi, blocks)
    vw = viterbi_wrapper(a, b, pi, blocks)
    for i in range(len(blocks)):
        assert np.array_equal(pw[i], out[f"post_{i}"])
        assert np.array_equal(vw[i], out[f"vit_{i}"])
    fn = os.path.join(HERE, "recursions_" + model_file.replace("model_", ""))
    np.savez_compressed(fn, **out)
    print(fn, "loglik_total", float(out["loglik_total"]), flush=True)


def main(argv):
    for job in argv:
        parts = job.split(":")
        if parts[0] == "symbols":
            job_symbols()
        elif parts[0] == "statespace":
            job_statespace()
        elif parts[0] == "model":
            job_model(int(parts[1]), int(parts[2]), parts[3])
        elif parts[0] == "recursions":
            job_recursions(parts[1])
        else:
            raise SystemExit(f"unknown job {job}")


if __name__ == "__main__":
    main(sys.argv[1:])
