"""Speculative Nelder-Mead: the simplex search the reference runs through
``scipy.optimize.minimize(method="Nelder-Mead", bounds=...)`` (optimizer.py:623-637),
restated so that every iteration costs ONE batched objective call instead of one to two
dependent ones.

The sequential algorithm evaluates the reflection point and then, depending on the value it
gets, an expansion, an outside contraction or an inside contraction.  All four candidates
are known before any of them is evaluated, and on the device a handful of parameter sets
advance through the forward sweep in the time of one (the sweep is bound by per-column
latency, not throughput: DESIGN 4.1).  So the four candidates go out as one batch, and the
decision tree then picks the values it needs.  A shrink step (N points) and the initial
simplex (N+1 points) are single batches as well.

Only the evaluations the sequential algorithm would have made are *consumed* (counted in
``nfev``, passed to ``consume`` in the sequential order), so the iterates, the returned
optimum, ``nit``/``nfev`` and the optimisation history are those of the sequential search
on the same objective values.  Coefficients and termination follow the standard
(non-adaptive) method: reflection 1, expansion 2, contraction 0.5, shrink 0.5;
``xatol = fatol = 1e-4``; initial simplex = x0 with each coordinate in turn enlarged by 5 %
(0.00025 where it is zero), reflected at the upper bound and clipped to the box.
"""
import warnings

import numpy as np
from scipy.optimize import OptimizeResult

RHO, CHI, PSI, SIGMA = 1.0, 2.0, 0.5, 0.5
NONZDELT, ZDELT = 0.05, 0.00025


class _Budget(Exception):
    pass


def initial_simplex(x0, lower=None, upper=None):
    x0 = np.asarray(x0, dtype=np.float64).ravel()
    if lower is not None:
        x0 = np.clip(x0, lower, upper)
    n = len(x0)
    sim = np.tile(x0, (n + 1, 1))
    for k in range(n):
        sim[k + 1, k] = (1 + NONZDELT) * x0[k] if x0[k] != 0 else ZDELT
    if lower is not None:
        sim = np.where(sim > upper, 2 * upper - sim, sim)
        sim = np.clip(sim, lower, upper)
    return sim


def minimize_neldermead_batched(batch_fun, x0, bounds=None, maxiter=None, maxfev=None,
                                xatol=1e-4, fatol=1e-4, consume=None, disp=False):
    """Minimise ``batch_fun`` (rows of an (m, N) array -> m values) from ``x0``.

    ``bounds``: sequence of (min, max) per coordinate, or None.  ``consume(x, f)`` is called
    once per evaluation the sequential method would have made, in its order.  Returns an
    ``OptimizeResult`` with scipy's fields plus ``nbatch`` (objective calls = device round
    trips) and ``nspec`` (points evaluated, speculative ones included)."""
    lower = upper = None
    if bounds is not None:
        lower = np.array([-np.inf if b[0] is None else b[0] for b in bounds], dtype=np.float64)
        upper = np.array([np.inf if b[1] is None else b[1] for b in bounds], dtype=np.float64)
        if (lower > upper).any():
            raise ValueError("Nelder Mead - one of the lower bounds is greater than an upper bound.")
    sim = initial_simplex(x0, lower, upper)
    n = sim.shape[1]
    if maxiter is None and maxfev is None:
        maxiter = maxfev = n * 200
    elif maxiter is None:
        maxiter = n * 200 if maxfev == np.inf else np.inf
    elif maxfev is None:
        maxfev = n * 200 if maxiter == np.inf else np.inf

    count = {"nfev": 0, "nbatch": 0, "nspec": 0}

    def evaluate(points):
        points = np.atleast_2d(points)
        vals = np.asarray(batch_fun(points), dtype=np.float64).ravel()
        if vals.shape[0] != points.shape[0]:
            raise ValueError("batch_fun must return one value per row")
        count["nbatch"] += 1
        count["nspec"] += points.shape[0]
        return vals

    def take(x, f):
        if count["nfev"] >= maxfev:
            raise _Budget
        count["nfev"] += 1
        if consume is not None:
            consume(x, f)
        return f

    def box(x):
        return x if lower is None else np.clip(x, lower, upper)

    fsim = np.full(n + 1, np.inf)
    first = n + 1 if maxfev >= n + 1 else int(maxfev)
    vals = evaluate(sim[:first]) if first > 0 else np.empty(0)
    try:
        for k in range(len(vals)):
            fsim[k] = take(sim[k], vals[k])
    except _Budget:
        pass
    order = np.argsort(fsim)
    sim, fsim = sim[order], fsim[order]

    iterations = 1
    while count["nfev"] < maxfev and iterations < maxiter:
        if (np.max(np.abs(sim[1:] - sim[0])) <= xatol and np.max(np.abs(fsim[0] - fsim[1:])) <= fatol):
            break
        try:
            xbar = np.add.reduce(sim[:-1], 0) / n
            worst = sim[-1]
            cand = np.stack([
                box((1 + RHO) * xbar - RHO * worst),                 # reflection
                box((1 + RHO * CHI) * xbar - RHO * CHI * worst),     # expansion
                box((1 + PSI * RHO) * xbar - PSI * RHO * worst),     # outside contraction
                box((1 - PSI) * xbar + PSI * worst),                 # inside contraction
            ])
            f = evaluate(cand)
            xr, xe, xc, xcc = cand
            fxr = take(xr, f[0])
            shrink = False
            if fxr < fsim[0]:
                fxe = take(xe, f[1])
                if fxe < fxr:
                    sim[-1], fsim[-1] = xe, fxe
                else:
                    sim[-1], fsim[-1] = xr, fxr
            elif fxr < fsim[-2]:
                sim[-1], fsim[-1] = xr, fxr
            elif fxr < fsim[-1]:
                fxc = take(xc, f[2])
                if fxc <= fxr:
                    sim[-1], fsim[-1] = xc, fxc
                else:
                    shrink = True
            else:
                fxcc = take(xcc, f[3])
                if fxcc < fsim[-1]:
                    sim[-1], fsim[-1] = xcc, fxcc
                else:
                    shrink = True
            if shrink:
                for j in range(1, n + 1):
                    sim[j] = box(sim[0] + SIGMA * (sim[j] - sim[0]))
                room = int(min(n, maxfev - count["nfev"]))
                fs = evaluate(sim[1:1 + room]) if room > 0 else np.empty(0)
                for j in range(1, n + 1):
                    if j - 1 >= len(fs):
                        raise _Budget
                    fsim[j] = take(sim[j], fs[j - 1])
            iterations += 1
        except _Budget:
            pass
        order = np.argsort(fsim)
        sim, fsim = sim[order], fsim[order]

    fval = float(np.min(fsim))
    status = 0
    if count["nfev"] >= maxfev:
        status, msg = 1, "Maximum number of function evaluations has been exceeded."
        if disp:
            warnings.warn(msg, RuntimeWarning, stacklevel=2)
    elif iterations >= maxiter:
        status, msg = 2, "Maximum number of iterations has been exceeded."
        if disp:
            warnings.warn(msg, RuntimeWarning, stacklevel=2)
    else:
        msg = "Optimization terminated successfully."
        if disp:
            print(msg)
            print(f"         Current function value: {fval:f}")
            print(f"         Iterations: {iterations:d}")
            print(f"         Function evaluations: {count['nfev']:d}")
            print(f"         Batched objective calls: {count['nbatch']:d} ({count['nspec']:d} points)")
    return OptimizeResult(fun=fval, nit=iterations, nfev=count["nfev"], status=status,
                          success=(status == 0), message=msg, x=sim[0], final_simplex=(sim, fsim),
                          nbatch=count["nbatch"], nspec=count["nspec"])
