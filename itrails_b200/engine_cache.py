"""Process-wide engines (one per GPU this process drives) and resident-block cache, so
that the reference-style wrappers (which receive ``V_lst`` on every call) upload the
alignment to HBM once instead of once per objective evaluation."""
from __future__ import annotations

from . import ngpu
from .engine import Engine

_ENGINE = None          # engine of the first local device (kept under this name for tests that swap it)
_ENGINES = {}           # device ordinal -> Engine (further local devices)
_LOADED = {}            # device ordinal -> (block list object, fingerprint) resident there


def get_engine(device=None):
    """The engine of ``device`` (default: this process's first GPU), created on first use."""
    global _ENGINE
    first = ngpu.local_devices()[0]
    if device is None or device == first:
        if _ENGINE is None:
            _ENGINE = Engine(first)
        return _ENGINE
    eng = _ENGINES.get(device)
    if eng is None:
        eng = _ENGINES[device] = Engine(device)
    return eng


def reset():
    global _ENGINE
    for eng in [_ENGINE, *_ENGINES.values()]:
        if eng is not None:
            eng.close()
    _ENGINE = None
    _ENGINES.clear()
    if _LOADED:
        _LOADED.clear()


def _fingerprint(V_lst):
    n = len(V_lst)
    probe = (0, n // 2, n - 1) if n else ()
    return (n, sum(len(v) for v in V_lst),
            tuple((id(V_lst[i]), len(V_lst[i]), int(V_lst[i][0]), int(V_lst[i][-1])) for i in probe))


def ensure_blocks(V_lst, device=None):
    """Make ``V_lst`` the resident data of the device's engine (no-op if it already is)."""
    global _LOADED
    if _LOADED is None:                     # (tests reset the cache by assigning None)
        _LOADED = {}
    eng = get_engine(device)
    key = getattr(eng, "device", 0)
    fp = _fingerprint(V_lst)
    hit = _LOADED.get(key)
    if hit is not None and hit[0] is V_lst and hit[1] == fp:
        return eng
    eng.load_blocks(V_lst)
    _LOADED[key] = (V_lst, fp)
    return eng
