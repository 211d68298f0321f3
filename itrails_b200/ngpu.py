"""Runtime device selection — the GPU analogue of the reference's ncpu.py:7-34.

The reference sizes a joblib pool from ``n_cpu``; here the blocks of an alignment are
spread over GPUs instead.  Two ways to use more than one GPU, both block-sharded with no
data-path collective:

* one process per GPU under ``torchrun`` (``LOCAL_RANK`` picks the device), or
* ONE process driving several GPUs (``update_n_gpu(n)``, the ``n_gpu`` setting of the
  YAML / ``--n_gpu`` flag of the decode CLIs, or ``ITRAILS_NGPU=n``): one context per
  device, calls dispatched from host threads (ctypes releases the GIL).

``update_n_cpu`` is kept as a shim so reference-style workflows keep working."""
import os

N_CPU_GLOBAL = 1
N_GPU_GLOBAL = None     # None: ITRAILS_NGPU or 1


def update_n_cpu(user_requested):
    """Accepted for compatibility with ncpu.update_n_cpu; host threads do not matter
    on this path (all numerics run on the GPU)."""
    global N_CPU_GLOBAL
    try:
        N_CPU_GLOBAL = max(1, int(user_requested))
    except (TypeError, ValueError):
        N_CPU_GLOBAL = os.cpu_count() or 1
    return N_CPU_GLOBAL


def update_n_gpu(user_requested):
    """GPUs this process drives (ignored under torchrun, where a process owns one GPU)."""
    global N_GPU_GLOBAL
    N_GPU_GLOBAL = None if user_requested is None else max(1, int(user_requested))
    return N_GPU_GLOBAL


def local_device():
    for key in ("ITRAILS_DEVICE", "LOCAL_RANK"):
        if key in os.environ:
            try:
                return int(os.environ[key])
            except ValueError:
                pass
    return 0


def local_devices():
    """Device ordinals this process drives."""
    if "LOCAL_RANK" in os.environ or "ITRAILS_DEVICE" in os.environ:
        return [local_device()]
    n = N_GPU_GLOBAL
    if n is None:
        try:
            n = max(1, int(os.environ.get("ITRAILS_NGPU", "1")))
        except ValueError:
            n = 1
    return list(range(n))
