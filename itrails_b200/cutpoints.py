"""Interval cutpoints — same names and values as the reference's cutpoints.py:5-65,
restated in closed form (no scipy): truncexpon.ppf / expon.ppf."""
import numpy as np


def cutpoints_AB(n_int_AB, t_AB, coal_AB):
    """cutpoints.py:5-26: quantiles of Exp(coal_AB) truncated to [0, t_AB]."""
    q = np.arange(n_int_AB + 1) / n_int_AB
    scale = 1 / coal_AB
    b = t_AB / scale
    with np.errstate(divide="ignore"):
        cut = -np.log1p(q * np.expm1(-b)) * scale
    # scipy's ppf returns the upper end of the support for q == 1 (b * scale + loc), not the
    # closed form — which loses digits from b ~ 20 and overflows to inf at b >= 37
    cut[-1] = b * scale
    return cut


def cutpoints_ABC(n_int_ABC, coal_ABC):
    """cutpoints.py:29-45: quantiles of Exp(coal_ABC); the last cutpoint is +inf."""
    q = np.arange(n_int_ABC + 1) / n_int_ABC
    with np.errstate(divide="ignore"):
        return -np.log1p(-q) / coal_ABC


def get_times(cut, intervals):
    """cutpoints.py:48-65."""
    return [cut[intervals[i + 1]] - cut[intervals[i]] for i in range(len(intervals) - 1)]
