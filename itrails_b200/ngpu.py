"""Runtime device selection — the GPU analogue of the reference's ncpu.py:7-34.

The reference sizes a joblib pool from ``n_cpu``; here one process drives one GPU
(``LOCAL_RANK`` under torchrun, else ``ITRAILS_DEVICE``, else 0) and ``update_n_cpu``
is kept as a shim so reference-style workflows keep working."""
import os

N_CPU_GLOBAL = 1


def update_n_cpu(user_requested):
    """Accepted for compatibility with ncpu.update_n_cpu; host threads do not matter
    on this path (all numerics run on the GPU)."""
    global N_CPU_GLOBAL
    try:
        N_CPU_GLOBAL = max(1, int(user_requested))
    except (TypeError, ValueError):
        N_CPU_GLOBAL = os.cpu_count() or 1
    return N_CPU_GLOBAL


def local_device():
    for key in ("ITRAILS_DEVICE", "LOCAL_RANK"):
        if key in os.environ:
            try:
                return int(os.environ[key])
            except ValueError:
                pass
    return 0
