#!/usr/bin/env python
"""bench.py — throughput of the coalescent-HMM hot path (alignment columns / second).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm (oracle port)

Workload (BASELINE.json configs[1], "config2"): a synthetic 4-species alignment of
10 Mb in 100 MAF blocks (lengths uniform 50-150 kb), default discretisation
n_int_AB = n_int_ABC = 3 (K = 27 hidden states), example_config.yaml parameters,
1 % of columns with an N.  One *step* = one pass of the hot path over that alignment:
one objective evaluation of itrails-optimize (GPU model build + forward
log-likelihood) + Viterbi (with traceback) + posterior decoding of every column; the
three recursions run concurrently on their own CUDA streams.  `value` = columns /
second for the whole step with the alignment resident in HBM; `breakdown` gives each
recursion's own kernel time; `e2e` is the same step through the C ABI with host
buffers (H2D of the symbols, model and tables, D2H of the log-likelihood, the Viterbi
path and the posterior matrix inside the timed region).
Under torchrun each rank owns its own 10 Mb alignment (weak scaling; the blocks of a
chromosome shard with no data-path collective) and the per-rank log-likelihoods are
summed with one NCCL all-reduce per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# read when the CUDA context is created (torch may create it first): see itrails_b200/__init__.py
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

SEED0 = 20261018
WORKLOADS = {
    # name: (config index, n_blocks, total columns, n_int_AB, n_int_ABC)
    "config1": (0, 1, 100_000, 3, 3),
    "config2": (1, 100, 10_000_000, 3, 3),
    "config3": (2, 1000, 100_000_000, 5, 5),
    "config4": (3, 2500, 250_000_000, 3, 3),
}


# ---------------------------------------------------------------------------------
def sample_clocks(stop, out, gpu_index):
    """SM clock and throttle reasons DURING the timed region.  NVML (a few samples per
    step) when the bindings are importable, else the nvidia-smi query of the profiling recipe."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = (0x8, 0x40, 0x20, 0x4)     # hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap (nvml.h)
        n_ok = 0
        while True:
            sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            try:
                r = int(reasons_fn(h))
            except Exception:
                r = 0
            out.append([str(sm), str(mx), "0"] + ["Active" if r & b else "Not Active" for b in bits])
            n_ok += 1
            if stop.is_set():
                break
            stop.wait(0.004)
        return
    except Exception as e:
        print(f"[bench] NVML clock sampling unavailable ({type(e).__name__}: {e}); using nvidia-smi", file=sys.stderr)
    q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    while True:                      # (at least one sample, even for a very short timed region)
        try:
            r = subprocess.run(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={q}",
                                "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
            f = [x.strip() for x in r.stdout.strip().split(",")]
            if len(f) >= 7:
                out.append(f)
        except Exception:
            pass
        if stop.is_set():
            break
        stop.wait(0.2)


def summarise_clocks(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
    sm = sorted(float(s[0]) for s in samples)
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in samples)]
    return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(samples[0][1]), "reasons": reasons,
            "samples": len(samples)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


# ---------------------------------------------------------------------------------
def get_model(eng, n_ab, n_abc, model_npz=None):
    """(a, b, pi) for the example parameters.  Product path: the GPU model builder.
    A fixture file can be forced with --model-npz (used before the builder existed)."""
    from itrails_b200 import synth
    if model_npz:
        g = np.load(model_npz)
        return g["a"], g["b"], g["pi"], "fixture:" + os.path.basename(model_npz)
    args = synth.example_model_args(n_abc)
    a, b, pi, _hidden = eng.build_model(args[None, :], n_ab, n_abc)
    return a[0], b[0], pi[0], "itr_build_model"


def make_workload(name, a, b, pi, rank, scale=1.0):
    from itrails_b200 import synth
    idx, n_blocks, total, _, _ = WORKLOADS[name]
    n_blocks = max(1, int(round(n_blocks * scale)))
    total = max(n_blocks, int(round(total * scale)))
    rng = np.random.default_rng(SEED0 + idx + 1000 * rank)
    lens = synth.block_lengths(n_blocks, total, rng) if n_blocks > 1 else np.array([total])
    return synth.alignment(a, b, pi, lens, SEED0 + idx + 1000 * rank + 1)


# ---------------------------------------------------------------------------------
def cpu_port_step(a, b, pi, V_lst, threads):
    """One step of the reference algorithm's C port (oracle/) on the host: forward
    log-likelihood + Viterbi + posterior.  Returns seconds per part."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import hmm_oracle as ho
    import hmm_oracle_c as hoc
    E = ho.emission_table(b)
    t0 = time.perf_counter()
    hoc.loglik_blocks(a, E, pi, V_lst, threads)
    t1 = time.perf_counter()
    LA, LE, om0 = ho.viterbi_tables(a, b, pi, V_lst)
    hoc.viterbi_blocks(LA, LE, om0, V_lst, threads)
    t2 = time.perf_counter()
    hoc.post_prob_blocks(a, E, pi, V_lst, threads)
    t3 = time.perf_counter()
    return t1 - t0, t2 - t1, t3 - t2


def cpu_sample(V_lst, max_cols):
    out, n = [], 0
    for V in V_lst:
        if n >= max_cols:
            break
        take = V[: max_cols - n]
        out.append(take)
        n += len(take)
    return out, n


# ---------------------------------------------------------------------------------
def _emit(line):
    """Print the one JSON line on the real stdout (see main: fd 1 is pointed at stderr while
    the benchmark runs, because NCCL and other native libraries write banners to it)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debug only; invalidates the number)")
    ap.add_argument("--model-npz", default=None)
    ap.add_argument("--cpu-sample-cols", type=int, default=2_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    _, _, _, n_ab, n_abc = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1

    # ---------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        g = np.load(args.model_npz or os.path.join(ROOT, "tests", "golden", "model_3_3_example.npz"))
        a, b, pi = g["a"], g["b"], g["pi"]
        V_all = make_workload(args.workload, a, b, pi, 0, args.scale)
        V_lst, ncol = cpu_sample(V_all, args.cpu_sample_cols)
        for _ in range(max(1, min(args.warmup, 1))):
            cpu_port_step(a, b, pi, V_lst[:2], cores)
        t0 = time.perf_counter()
        parts = np.zeros(3)
        for _ in range(args.steps):
            parts += cpu_port_step(a, b, pi, V_lst, cores)
        dt = (time.perf_counter() - t0) / args.steps
        val = ncol / dt
        line = {
            "impl": "reference", "metric": "alignment columns/sec (forward loglik + Viterbi + posterior)",
            "value": val, "unit": "columns/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "K": int(a.shape[0]), "sample_columns": ncol,
                       "blocks": len(V_lst)},
            "cpu_baseline": {"value": val, "unit": "columns/s", "cores": cores, "kind": "port",
                             "sample": f"first {ncol} columns ({len(V_lst)} blocks) of {args.workload}, "
                                       "C port of optimizer.py:146-354 (oracle/hmm_oracle.c), all host threads over blocks"},
            "breakdown": {k: ncol * args.steps / max(v, 1e-12) for k, v in zip(("forward", "viterbi", "posterior"), parts)},
            "e2e": {"value": val, "unit": "columns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        _emit(line)
        return 0

    # ---------------------------------------------------------------- our arm
    import torch
    import itrails_b200 as itb
    from itrails_b200 import distributed as D
    from itrails_b200 import synth
    from itrails_b200.optimizer import viterbi_tables

    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(local_rank)
    eng = itb.Engine(local_rank)
    info = eng.device_info()
    a, b, pi, model_src = get_model(eng, n_ab, n_abc, args.model_npz)
    params = synth.example_model_args(n_abc)[None, :]
    K = a.shape[0]
    V_lst = make_workload(args.workload, a, b, pi, rank, args.scale)
    sym, off = itb.Engine.pack_blocks(V_lst)
    ncol = int(off[-1])
    log_a, log_E, omega0 = viterbi_tables(a, b, pi, V_lst)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # resident step -----------------------------------------------------------------
    # One objective evaluation (GPU model build + forward log-likelihood, what
    # itrails-optimize does per iteration) + Viterbi + posterior decoding of the
    # resident alignment.  The three recursions are enqueued on their own streams and
    # overlap on the device (itr_set_async); the step ends at itr_sync.
    eng.load_packed(sym, off)
    rebuild = args.model_npz is None

    def step_resident():
        eng.set_async(True)
        if rebuild:
            eng.build_model(params, n_ab, n_abc, fetch=False)     # enqueued; Viterbi (own tables) overlaps it
        eng.viterbi(log_a, log_E, omega0, fetch=False)
        eng.posterior(fetch=False)
        ll = eng.loglik()
        eng.sync()
        eng.set_async(False)
        if world > 1:
            ll = D.allreduce_sum(ll, local_rank)
        return ll

    if not rebuild:
        eng.set_model(a, b, pi)
    for _ in range(args.warmup):
        step_resident()
    clk_samples, stop = [], threading.Event()
    th = threading.Thread(target=sample_clocks, args=(stop, clk_samples, local_rank), daemon=True)
    if rank == 0:
        th.start()
    phases = ("model", "loglik", "viterbi_fwd", "viterbi_trace", "post_fwd", "post_bwd", "post_combine", "post_total")
    ph_ms = {p: 0.0 for p in phases}
    barrier()
    launches_before = eng.launch_count
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ll = step_resident()
        for p in phases:
            ph_ms[p] += max(eng.phase_ms(p), 0.0)
    barrier()
    dt = time.perf_counter() - t0
    launches = eng.launch_count - launches_before
    stop.set()
    dt = D.allreduce_max(dt)
    ms_step = dt * 1e3 / args.steps
    for p in phases:
        ph_ms[p] /= args.steps
    # critical path on the device: Viterbi (own log tables) overlaps the model build, the other two wait for it
    dev_ms = max(ph_ms["viterbi_fwd"] + ph_ms["viterbi_trace"], ph_ms["model"] + max(ph_ms["loglik"], ph_ms["post_total"]))
    total_cols = float(D.allreduce_sum(np.array([float(ncol)]), local_rank)[0])
    value = total_cols / (ms_step * 1e-3)

    # e2e step: host buffers in, host buffers out ------------------------------------
    # Through the C ABI with page-locked host buffers: upload of the symbols, block
    # offsets, model matrices and Viterbi tables; download of the log-likelihood, the
    # state path and the full posterior matrix, all inside the timed region.
    e2e = None
    if not args.no_e2e:
        pin = lambda n, dt_: torch.empty(n, dtype=dt_, pin_memory=True).numpy()
        sym_pin = pin(len(sym), torch.uint16)
        sym_pin[:] = sym
        path_pin = pin(ncol, torch.uint8)
        post_pin = pin(ncol * K, torch.float64).reshape(ncol, K)

        def step_e2e():
            eng.load_packed(sym_pin, off)
            eng.set_model(a, b, pi)
            eng.set_async(True)
            eng.posterior(out=post_pin)          # largest download first: it overlaps the rest
            eng.viterbi(log_a, log_E, omega0, out=path_pin)
            ll = eng.loglik()
            eng.sync()
            eng.set_async(False)
            if world > 1:
                ll = D.allreduce_sum(ll, local_rank)
            return ll

        step_e2e()
        step_e2e()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, args.steps)
        e2e_ms = []
        for _ in range(n_e2e):
            t1 = time.perf_counter()
            step_e2e()
            e2e_ms.append((time.perf_counter() - t1) * 1e3)
        barrier()
        dte = D.allreduce_max((time.perf_counter() - t0) / n_e2e)
        h2d = sym.nbytes + off.nbytes + (a.nbytes + b.nbytes + pi.nbytes) + log_a.nbytes + log_E.nbytes + omega0.nbytes
        d2h = ncol * 1 + ncol * K * 8 + 8 * len(V_lst)
        e2e = {"value": total_cols / dte, "unit": "columns/s", "ms_per_step": dte * 1e3,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": n_e2e,
               "per_step_ms": [round(x, 2) for x in e2e_ms]}
        eng.load_packed(sym, off)
        eng.set_model(a, b, pi)

    if rank != 0:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
        return 0

    # roofline of the dominant kernel ---------------------------------------------------
    hbm_peak, peak_src = measured_peaks()
    dom = max(("loglik", "viterbi_fwd", "viterbi_trace", "post_fwd", "post_bwd", "post_combine"), key=lambda p: ph_ms[p])
    alg_bytes = {  # per column, algorithmic
        "loglik": 2, "viterbi_fwd": 2 + 32, "viterbi_trace": 32 + 1, "post_fwd": 2 + 8 * K, "post_bwd": 2 + 8 * K,
        "post_combine": 24 * K}
    alg_flops = {"loglik": 2 * K * K + 3 * K, "viterbi_fwd": 3 * K * K, "viterbi_trace": 0,
                 "post_fwd": 2 * K * K + 3 * K, "post_bwd": 2 * K * K + 3 * K, "post_combine": 3 * K}
    dom_s = ph_ms[dom] * 1e-3
    prof = {}
    pf = os.path.join(ROOT, "profiles", "measured_r1.json")
    if os.path.exists(pf):
        prof = json.load(open(pf))
    fp64_peak = prof.get("fp64_tflops")            # DFMA/DMMA issue peak measured by tools/ubench.cu
    traffic = (prof.get("dram_bytes_per_launch") or {}).get(dom) if args.workload == "config2" and args.scale == 1.0 else None
    roof = {"kernel": dom, "bound": "hbm", "achieved": alg_bytes[dom] * ncol / dom_s / 1e9, "peak": hbm_peak,
            "unit": "GB/s", "peak_source": peak_src, "traffic": traffic,
            "algorithmic_bytes_per_launch": alg_bytes[dom] * ncol}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["fp64"] = {"achieved": alg_flops[dom] * ncol / dom_s / 1e12, "unit": "TFLOP/s", "peak": fp64_peak,
                    "frac": (alg_flops[dom] * ncol / dom_s / 1e12 / fp64_peak) if fp64_peak else None,
                    "note": "the recursions are dependent chains of FP64 add/compare (one chain per alignment block); "
                            "with 100 blocks they are bound by per-column latency, not by HBM or FP64 throughput "
                            "(DESIGN.md, SURVEY 8d)"}

    # CPU baseline (oracle port) on a bounded sample ---------------------------------------
    cpu = None
    if not args.no_cpu_baseline and world >= 1:
        V_s, ncs = cpu_sample(V_lst, args.cpu_sample_cols)
        parts = cpu_port_step(a, b, pi, V_s, cores)
        cpu = {"value": ncs / sum(parts), "unit": "columns/s", "cores": cores, "kind": "port",
               "sample": f"first {ncs} columns ({len(V_s)} blocks) of the workload; C port of "
                         "optimizer.py:146-354 (oracle/hmm_oracle.c), one pass, all host threads over blocks",
               "breakdown": {k: ncs / max(v, 1e-12) for k, v in zip(("forward", "viterbi", "posterior"), parts)}}

    if th.is_alive():
        th.join(timeout=10)
    line = {
        "metric": "alignment columns/sec (forward loglik + Viterbi + posterior)",
        "value": value, "unit": "columns/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "device_ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload if args.scale == 1.0 else f"{args.workload}*{args.scale}",
                   "columns_per_gpu": ncol, "blocks_per_gpu": len(V_lst), "K": int(K),
                   "n_int_AB": n_ab, "n_int_ABC": n_abc, "model": model_src,
                   "l2": "inputs+outputs per step (%.0f MB) exceed L2 (126 MB)" % ((ncol * (2 + 32 + 8 * K)) / 1e6),
                   "device": info["name"], "sm_count": info["sm_count"]},
        "breakdown": {"forward": ncol / (ph_ms["loglik"] * 1e-3),
                      "viterbi": ncol / ((ph_ms["viterbi_fwd"] + ph_ms["viterbi_trace"]) * 1e-3),
                      "posterior": ncol / (ph_ms["post_total"] * 1e-3),
                      "unit": "columns/s per GPU, kernel time only", "phase_ms": ph_ms},
        "loglik": float(np.atleast_1d(ll)[0]),
        "gpu_launches": int(launches),
        "clocks": summarise_clocks(clk_samples),
        "roofline": roof,
        "cpu_baseline": cpu,
        "e2e": e2e,
    }
    _emit(line)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
