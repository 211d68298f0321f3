#!/usr/bin/env python
"""bench.py — throughput of the coalescent-HMM hot path (alignment columns / second).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm (oracle port)

Workload (BASELINE.json configs[3], "config4" — the configuration the metric is quoted
on): ONE chromosome-scale synthetic 4-species alignment of 250 Mb in 2 500 MAF blocks
(lengths uniform 50-150 kb), default discretisation n_int_AB = n_int_ABC = 3 (K = 27
hidden states), example_config.yaml parameters, 1 % of columns with an N.  With N GPUs
the blocks of that one alignment are LPT-partitioned over the ranks (the product's own
partition, itrails_b200.distributed.lpt_partition) — STRONG scaling: the total work is
fixed, `value` = 250e6 columns / (slowest rank's time per step).  No data-path
collective; one scalar all-reduce of the log-likelihood per step.

One *step* = one pass of the hot path over the alignment: one objective evaluation of
itrails-optimize (GPU model build + forward log-likelihood) + Viterbi (with traceback) +
posterior decoding of every column; the three recursions are enqueued on their own CUDA
streams.  `value`: alignment resident in HBM, results left in HBM.  `e2e`: the same step
through the C ABI with HOST buffers — symbols, model and tables uploaded from pinned
memory, log-likelihood and state path downloaded, and the FP64 posterior matrix (54 GB)
streamed to the host through a bounded ring of pinned buffers (itr_posterior_stream),
every byte counted.  `extra` (N = 1 only): config 2 (10 Mb / 100 blocks — the
latency-bound case), objective evaluations / s and the model build, each with the CPU
port's time beside it.
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
FP64_PEAK_FALLBACK = 37.06      # TFLOP/s, DMMA m8n8k4 whole-GPU issue peak measured by tools/ubench.cu (profiles/ubench_r1.txt)


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
        while True:
            sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            try:
                r = int(reasons_fn(h))
            except Exception:
                r = 0
            out.append([str(sm), str(mx), "0"] + ["Active" if r & b else "Not Active" for b in bits])
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
    """(HBM GB/s, FP64 TFLOP/s, source): the driver-written copy peak, and this repo's own
    micro-benchmark of the FP64 pipe (MEASURED_PEAKS.json has no FP64 figure)."""
    hbm, src = 6650.0, "fallback"
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        hbm, src = float(json.load(open(p))["hbm_gbs"]), "measured"
    fp64 = FP64_PEAK_FALLBACK
    for name in ("measured_r2.json", "measured_r1.json"):
        pf = os.path.join(ROOT, "profiles", name)
        if os.path.exists(pf):
            fp64 = float(json.load(open(pf)).get("fp64_tflops") or fp64)
            break
    return hbm, fp64, src


# ---------------------------------------------------------------------------------
def workload_lengths(name, scale=1.0):
    from itrails_b200 import synth
    idx, n_blocks, total, _, _ = WORKLOADS[name]
    n_blocks = max(1, int(round(n_blocks * scale)))
    total = max(n_blocks, int(round(total * scale)))
    rng = np.random.default_rng(SEED0 + idx)
    return synth.block_lengths(n_blocks, total, rng) if n_blocks > 1 else np.array([total], dtype=np.int64)


def workload_blocks(name, a, b, pi, lengths, ids):
    """Blocks `ids` of the workload's alignment (uint16 symbols; block i has its own random
    stream, so every rank generates exactly its share of the ONE alignment)."""
    from itrails_b200 import synth
    return synth.alignment_blocks(a, b, pi, lengths, ids, SEED0 + 100 + WORKLOADS[name][0], dtype=np.uint16)


def pack(V_lst):
    off = np.zeros(len(V_lst) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(v) for v in V_lst])
    return np.ascontiguousarray(np.concatenate(V_lst), dtype=np.uint16), off


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


def cpu_sample_ids(lengths, cores, max_cols):
    """At least four blocks per host thread (so that no thread idles on the tail), taken
    from the start of the alignment, capped at `max_cols` columns."""
    want = min(len(lengths), max(4 * cores, 8))
    ids, n = [], 0
    for i in range(want):
        if ids and n + int(lengths[i]) > max_cols:
            break
        ids.append(i)
        n += int(lengths[i])
    return ids, n


# ---------------------------------------------------------------------------------
def _emit(line):
    """Print the one JSON line on the real stdout (see main: fd 1 is pointed at stderr while
    the benchmark runs, because NCCL and other native libraries write banners to it)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1
PHASES = ("model", "loglik", "viterbi_fwd", "viterbi_trace", "post_fwd", "post_bwd", "post_combine", "post_total")


class Runner:
    """The resident and the end-to-end step of one workload on one engine."""

    def __init__(self, eng, params, n_ab, n_abc, a, b, pi, V_lst, world, local_rank):
        import torch
        from itrails_b200.optimizer import viterbi_tables
        self.torch = torch
        self.eng, self.params, self.n_ab, self.n_abc = eng, params, n_ab, n_abc
        self.a, self.b, self.pi = a, b, pi
        self.world, self.local_rank = world, local_rank
        self.K = a.shape[0]
        self.sym, self.off = pack(V_lst)
        self.ncol = int(self.off[-1])
        self.nblk = len(V_lst)
        self.log_a, self.log_E, self.omega0 = viterbi_tables(a, b, pi, V_lst)
        self.pins = None

    def reduce(self, ll):
        from itrails_b200 import distributed as D
        return D.allreduce_sum(ll, self.local_rank) if self.world > 1 else ll

    # resident step: model build + loglik || Viterbi || posterior, everything stays in HBM
    def load(self):
        self.eng.load_packed(self.sym, self.off)

    def step_resident(self):
        eng = self.eng
        eng.set_async(True)
        eng.build_model(self.params, self.n_ab, self.n_abc, fetch=False)     # enqueued; Viterbi (own tables) overlaps it
        eng.viterbi(self.log_a, self.log_E, self.omega0, fetch=False)
        eng.posterior(fetch=False)
        ll = eng.loglik()
        eng.sync()
        eng.set_async(False)
        return self.reduce(ll)

    # e2e step: host buffers in, host buffers out
    def prepare_e2e(self, ring_bytes=1 << 30, n_slots=4, stream=True):
        torch = self.torch
        pin = lambda n, dt_: torch.empty(n, dtype=dt_, pin_memory=True).numpy()   # noqa: E731  (cudaHostAlloc, once)
        sym_pin = pin(len(self.sym), torch.uint16)
        sym_pin[:] = self.sym
        path_pin = pin(self.ncol, torch.uint8)
        if stream:
            slot_cols = max(1, min(self.ncol, ring_bytes // n_slots // (8 * self.K)))
            ring = pin(n_slots * slot_cols * self.K, torch.float64)
            self.pins = (sym_pin, path_pin, ring, slot_cols, n_slots)
        else:
            post_pin = pin(self.ncol * self.K, torch.float64).reshape(self.ncol, self.K)
            self.pins = (sym_pin, path_pin, post_pin, 0, 0)
        self.stream = stream

    def step_e2e(self):
        eng = self.eng
        sym_pin, path_pin, dst, slot_cols, n_slots = self.pins
        eng.load_packed(sym_pin, self.off)
        eng.set_model(self.a, self.b, self.pi)
        eng.set_async(True)
        if self.stream:
            eng.posterior(fetch=False)                         # enqueued first: its rows are what the PCIe link waits for
            eng.viterbi(self.log_a, self.log_E, self.omega0, out=path_pin)
            ll = eng.loglik()
            eng.posterior_stream(dst, slot_cols, n_slots)      # drains it; returns when the last piece is in host memory
        else:
            eng.posterior(out=dst)          # largest download first: it overlaps the rest
            eng.viterbi(self.log_a, self.log_E, self.omega0, out=path_pin)
            ll = eng.loglik()
        eng.sync()
        eng.set_async(False)
        return self.reduce(ll)

    def e2e_bytes(self):
        h2d = self.sym.nbytes + self.off.nbytes + (self.a.nbytes + self.b.nbytes + self.pi.nbytes) + \
            self.log_a.nbytes + self.log_E.nbytes + self.omega0.nbytes
        d2h = self.ncol * 1 + self.ncol * self.K * 8 + 8 * self.nblk
        return int(h2d), int(d2h)


def timed(fn, steps, barrier):
    barrier()
    t0 = time.perf_counter()
    out = None
    for _ in range(steps):
        out = fn()
    barrier()
    return (time.perf_counter() - t0) / steps, out


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
    ap.add_argument("--workload", default="config4", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debug only; invalidates the number)")
    ap.add_argument("--cpu-sample-cols", type=int, default=16_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    _, _, _, n_ab, n_abc = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    wl_name = args.workload if args.scale == 1.0 else f"{args.workload}*{args.scale}"
    lengths = workload_lengths(args.workload, args.scale)
    total_cols = int(lengths.sum())

    # ---------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        g = np.load(os.path.join(ROOT, "tests", "golden", f"model_{n_ab}_{n_abc}_example.npz")) \
            if (n_ab, n_abc) == (3, 3) else None
        if g is None:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import ctmc_oracle
            from itrails_b200 import synth
            a, b, pi = ctmc_oracle.trans_emiss_calc(*synth.example_model_args(n_abc), n_ab, n_abc)[:3]
        else:
            a, b, pi = g["a"], g["b"], g["pi"]
        ids, ncol = cpu_sample_ids(lengths, cores, args.cpu_sample_cols)
        V_lst = [v.astype(np.int64) for v in workload_blocks(args.workload, a, b, pi, lengths, ids)]
        for _ in range(max(1, min(args.warmup, 1))):
            cpu_port_step(a, b, pi, V_lst[:max(2, cores // 4)], cores)
        t0 = time.perf_counter()
        parts = np.zeros(3)
        for _ in range(args.steps):
            parts += cpu_port_step(a, b, pi, V_lst, cores)
        dt = (time.perf_counter() - t0) / args.steps
        val = ncol / dt
        line = {
            "impl": "reference", "metric": "alignment columns/sec (forward loglik + Viterbi + posterior)",
            "value": val, "unit": "columns/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl_name, "K": int(a.shape[0]), "sample_columns": ncol, "blocks": len(V_lst)},
            "cpu_baseline": {"value": val, "unit": "columns/s", "cores": cores, "kind": "port",
                             "sample": f"first {len(V_lst)} blocks ({ncol} columns, >= 4 blocks per host thread) of {wl_name}; "
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

    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = itb.Engine(local_rank)
    info = eng.device_info()
    params = synth.example_model_args(n_abc)[None, :]
    a, b, pi, _hidden = eng.build_model(params, n_ab, n_abc)          # product path: the GPU model builder
    a, b, pi = a[0], b[0], pi[0]
    K = a.shape[0]

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # this rank's share of the ONE alignment: the product's LPT partition
    parts = D.lpt_partition(lengths, world)
    my_ids = parts[rank]
    per_rank_cols = [int(lengths[p].sum()) for p in parts]
    t_gen = time.perf_counter()
    V_lst = workload_blocks(args.workload, a, b, pi, lengths, my_ids)
    t_gen = time.perf_counter() - t_gen
    run = Runner(eng, params, n_ab, n_abc, a, b, pi, V_lst, world, local_rank)
    ncol = run.ncol
    print(f"[bench] rank {rank}: {len(my_ids)} blocks, {ncol} columns (generated in {t_gen:.1f} s)", file=sys.stderr)

    run.load()
    for _ in range(args.warmup):
        run.step_resident()
    clk_samples, stop = [], threading.Event()
    th = threading.Thread(target=sample_clocks, args=(stop, clk_samples, local_rank), daemon=True)
    if rank == 0:
        th.start()
    ph_ms = {p: 0.0 for p in PHASES}
    barrier()
    launches_before = eng.launch_count
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ll = run.step_resident()
        for p in PHASES:
            ph_ms[p] += max(eng.phase_ms(p), 0.0)
    barrier()
    dt = time.perf_counter() - t0
    launches = eng.launch_count - launches_before
    stop.set()
    dt = D.allreduce_max(dt)
    ms_step = dt * 1e3 / args.steps
    for p in PHASES:
        ph_ms[p] /= args.steps
    value = total_cols / (ms_step * 1e-3)

    # every recursion alone (nothing else on the device): the kernel times the roofline uses
    alone = {}
    for name, fn in (("loglik", lambda: eng.loglik()), ("viterbi", lambda: eng.viterbi(run.log_a, run.log_E, run.omega0, fetch=False)),
                     ("posterior", lambda: eng.posterior(fetch=False))):
        fn()
        barrier()
        acc = {p: 0.0 for p in PHASES}
        for _ in range(3):
            fn()
            for p in PHASES:
                acc[p] += max(eng.phase_ms(p), 0.0) / 3
        alone[name] = acc
    alone_ms = {"loglik": alone["loglik"]["loglik"], "viterbi_fwd": alone["viterbi"]["viterbi_fwd"],
                "viterbi_trace": alone["viterbi"]["viterbi_trace"], "post_fwd": alone["posterior"]["post_fwd"],
                "post_bwd": alone["posterior"]["post_bwd"], "post_combine": alone["posterior"]["post_combine"],
                "post_total": alone["posterior"]["post_total"]}

    # e2e step ------------------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        run.prepare_e2e(stream=True)
        run.step_e2e()
        run.step_e2e()
        n_e2e = max(1, args.steps)
        e2e_ms = []

        def one():
            t1 = time.perf_counter()
            r = run.step_e2e()
            e2e_ms.append((time.perf_counter() - t1) * 1e3)
            return r
        dte, _ = timed(one, n_e2e, barrier)
        dte = D.allreduce_max(dte)
        h2d, d2h = run.e2e_bytes()
        h2d_all = float(D.allreduce_sum(np.array([float(h2d)]), local_rank)[0])
        d2h_all = float(D.allreduce_sum(np.array([float(d2h)]), local_rank)[0])
        e2e = {"value": total_cols / dte, "unit": "columns/s", "ms_per_step": dte * 1e3,
               "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h_all), "steps": n_e2e,
               "per_step_ms": [round(x, 2) for x in e2e_ms],
               "d2h_gbs_this_rank": d2h / dte / 1e9, "d2h_gbs_all_ranks": d2h_all / dte / 1e9,
               "posterior_path": "itr_posterior_stream: 4-slot pinned ring of 1 GiB, pieces <= %d columns" % run.pins[3]}
        ceil_file = os.path.join(ROOT, "profiles", "d2h_ceiling_r2.json")
        if os.path.exists(ceil_file):
            ceil = json.load(open(ceil_file)).get(str(world))
            if ceil:
                e2e["d2h_ceiling_gbs_all_ranks"] = ceil
                e2e["frac_of_d2h_ceiling"] = e2e["d2h_gbs_all_ranks"] / ceil
        run.pins = None
        run.load()

    imbalance = max(per_rank_cols) / (sum(per_rank_cols) / world)
    if rank != 0:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
        return 0

    # roofline of the dominant kernel ---------------------------------------------------
    hbm_peak, fp64_peak, peak_src = measured_peaks()
    kernels = ("loglik", "viterbi_fwd", "viterbi_trace", "post_fwd", "post_bwd", "post_combine")
    dom = max(kernels, key=lambda p: ph_ms[p])
    alg_bytes = {  # per column, algorithmic
        "loglik": 2, "viterbi_fwd": 2 + 32, "viterbi_trace": 32 + 1, "post_fwd": 2 + 8, "post_bwd": 2 + 8,
        "post_combine": 2 + 16 + 8 * K}
    alg_flops = {"loglik": 2 * K * K + 3 * K, "viterbi_fwd": 2 * K * K + K, "viterbi_trace": 0,
                 "post_fwd": 2 * K * K + 3 * K, "post_bwd": 2 * K * K + 3 * K,
                 "post_combine": 2 * (2 * K * K + 3 * K) + 3 * K}
    kernel_names = {"loglik": "forward_runs_kernel", "viterbi_fwd": "viterbi sweep", "viterbi_trace": "viterbi traceback",
                    "post_fwd": "checkpoint_sweep_kernel<fwd>", "post_bwd": "checkpoint_sweep_kernel<bwd>",
                    "post_combine": "posterior tiles kernel (pass 2)"}
    dom_s = ph_ms[dom] * 1e-3
    hbm_view = {"achieved": alg_bytes[dom] * ncol / dom_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": alg_bytes[dom] * ncol / dom_s / 1e9 / hbm_peak, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes[dom] * ncol}
    fp64_view = {"achieved": alg_flops[dom] * ncol / dom_s / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                 "frac": alg_flops[dom] * ncol / dom_s / 1e12 / fp64_peak,
                 "peak_source": "tools/ubench.cu: FP64 DMMA m8n8k4 whole-GPU issue rate (MEASURED_PEAKS.json has no FP64 figure)",
                 "algorithmic_flops_per_launch": alg_flops[dom] * ncol}
    prof = {}
    pf = os.path.join(ROOT, "profiles", "measured_r2.json")
    if os.path.exists(pf):
        prof = json.load(open(pf))
    traffic = (prof.get("dram_bytes_per_launch") or {}).get(f"{args.workload}:{dom}") if args.scale == 1.0 and world == 1 else None
    # pass 2 of the posterior is a dense FP64 contraction (2(2K^2+3K)+3K flop against 8K+18
    # bytes per column: 13.6 flop/B, above the 5.7 flop/B ridge) -> FP64 tensor pipe; every
    # other kernel is reported in the HBM form with the FP64 view beside it
    main_view = fp64_view if dom == "post_combine" else hbm_view
    roof = {"kernel": dom, "kernel_name": kernel_names[dom], "bound": "tensor" if dom == "post_combine" else "hbm",
            "achieved": main_view["achieved"], "peak": main_view["peak"], "unit": main_view["unit"],
            "frac": main_view["frac"], "traffic": traffic, "launch_ms": ph_ms[dom],
            "launch_ms_alone": alone_ms[dom],
            "frac_alone": main_view["frac"] * ph_ms[dom] / alone_ms[dom] if alone_ms[dom] > 0 else None,
            "hbm": hbm_view, "fp64": fp64_view,
            "note": "launch_ms: CUDA events around the kernel on its own stream inside the timed step (the other two "
                    "recursions share the SMs); launch_ms_alone: the same launch with nothing else on the device"}

    # CPU baseline (oracle port) on a bounded sample ---------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        ids, ncs = cpu_sample_ids(lengths, cores, args.cpu_sample_cols)
        V_s = [v.astype(np.int64) for v in workload_blocks(args.workload, a, b, pi, lengths, ids)]
        parts_s = cpu_port_step(a, b, pi, V_s, cores)
        cpu = {"value": ncs / sum(parts_s), "unit": "columns/s", "cores": cores, "kind": "port",
               "sample": f"first {len(V_s)} blocks ({ncs} columns, >= 4 blocks per host thread) of the workload; C port of "
                         "optimizer.py:146-354 (oracle/hmm_oracle.c), one pass, all host threads over blocks",
               "breakdown": {k: ncs / max(v, 1e-12) for k, v in zip(("forward", "viterbi", "posterior"), parts_s)}}
        del V_s

    # extras (one GPU): config 2, objective evaluations, model build --------------------------
    extra = None
    if not args.no_extra and world == 1 and args.scale == 1.0:
        extra = run_extras(eng, params, n_ab, n_abc, a, b, pi, cores, barrier, args)

    if th.is_alive():
        th.join(timeout=10)
    line = {
        "metric": "alignment columns/sec (forward loglik + Viterbi + posterior)",
        "value": value, "unit": "columns/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl_name, "total_columns": total_cols, "total_blocks": int(len(lengths)),
                   "columns_per_rank": per_rank_cols, "imbalance_max_over_mean": imbalance,
                   "partition": "LPT over blocks (itrails_b200.distributed.lpt_partition)", "K": int(K),
                   "n_int_AB": n_ab, "n_int_ABC": n_abc, "model": "itr_build_model",
                   "l2": "inputs+outputs per step per GPU (%.0f MB) exceed L2 (126 MB)" % ((ncol * (2 + 32 + 8 * K)) / 1e6),
                   "device": info["name"], "sm_count": info["sm_count"]},
        "breakdown": {"forward": ncol / (ph_ms["loglik"] * 1e-3),
                      "viterbi": ncol / ((ph_ms["viterbi_fwd"] + ph_ms["viterbi_trace"]) * 1e-3),
                      "posterior": ncol / (ph_ms["post_total"] * 1e-3),
                      "unit": "columns/s on rank 0, kernel time inside the step", "phase_ms": ph_ms,
                      "alone": {"forward": ncol / (alone_ms["loglik"] * 1e-3),
                                "viterbi": ncol / ((alone_ms["viterbi_fwd"] + alone_ms["viterbi_trace"]) * 1e-3),
                                "posterior": ncol / (alone_ms["post_total"] * 1e-3), "phase_ms": alone_ms}},
        "loglik": float(np.atleast_1d(ll)[0]),
        "gpu_launches": int(launches),
        "clocks": summarise_clocks(clk_samples),
        "roofline": roof,
        "cpu_baseline": cpu,
        "e2e": e2e,
        "extra": extra,
    }
    _emit(line)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_extras(eng, params, n_ab, n_abc, a, b, pi, cores, barrier, args):
    """Config 2 (the latency-bound case), the optimiser's objective and the model build on
    one GPU, each with the CPU port's time beside it."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ctmc_oracle
    import hmm_oracle as ho
    import hmm_oracle_c as hoc
    extra = {}
    lengths2 = workload_lengths("config2")
    V2 = workload_blocks("config2", a, b, pi, lengths2, range(len(lengths2)))
    run2 = Runner(eng, params, n_ab, n_abc, a, b, pi, V2, 1, eng.device)
    run2.load()
    for _ in range(3):
        run2.step_resident()
    steps2 = max(args.steps, 5)
    ph2 = {p: 0.0 for p in PHASES}

    def one():
        r = run2.step_resident()
        for p in PHASES:
            ph2[p] += max(eng.phase_ms(p), 0.0) / steps2
        return r
    dt2, _ = timed(one, steps2, barrier)
    c2 = {"workload": "config2", "columns": run2.ncol, "blocks": run2.nblk, "value": run2.ncol / dt2, "unit": "columns/s",
          "ms_per_step": dt2 * 1e3, "phase_ms": ph2}
    if not args.no_e2e:
        run2.prepare_e2e(stream=False)
        run2.step_e2e()
        run2.step_e2e()
        dte2, _ = timed(run2.step_e2e, steps2, barrier)
        h2d, d2h = run2.e2e_bytes()
        c2["e2e"] = {"value": run2.ncol / dte2, "unit": "columns/s", "ms_per_step": dte2 * 1e3,
                     "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "d2h_gbs": d2h / dte2 / 1e9}
        run2.pins = None
        run2.load()
    extra["config2"] = c2

    # objective evaluations of itrails-optimize on config 2: model build + forward log-likelihood,
    # one synchronous call pair per evaluation (optimization_wrapper, optimizer.py:396-583)
    rng = np.random.default_rng(3)
    n_eval = 50
    pts = params * (1.0 + 0.02 * rng.standard_normal((n_eval, params.shape[1])))

    def evals():
        for row in pts:
            eng.build_model(row[None, :], n_ab, n_abc, fetch=False)
            eng.loglik()
    evals()
    dte, _ = timed(evals, 1, barrier)
    # the model build alone: one parameter set, and a batch of 1024 (config 5's builder half)
    eng.build_model(params, n_ab, n_abc, fetch=False)
    build1 = []
    for _ in range(10):
        eng.build_model(params, n_ab, n_abc, fetch=False)
        build1.append(eng.phase_ms("model"))
    batch = params * (1.0 + 0.02 * rng.standard_normal((1024, params.shape[1])))
    eng.build_model(batch, n_ab, n_abc, fetch=False)
    eng.build_model(batch, n_ab, n_abc, fetch=False)
    build1024 = eng.phase_ms("model")
    t0 = time.perf_counter()
    ll1024 = eng.loglik()
    sweep_s = time.perf_counter() - t0
    eng.build_model(params, n_ab, n_abc, fetch=False)
    # the CPU port beside it: NumPy restatement of the model build, C port of the forward sweep
    t0 = time.perf_counter()
    ctmc_oracle.trans_emiss_calc(*params[0], n_ab, n_abc)
    cpu_build = time.perf_counter() - t0
    ids, ncs = cpu_sample_ids(lengths2, cores, 10_000_000)
    V_s = [V2[i].astype(np.int64) for i in ids]
    E = ho.emission_table(b)
    t0 = time.perf_counter()
    hoc.loglik_blocks(a, E, pi, V_s, cores)
    cpu_ll = (time.perf_counter() - t0) * (run2.ncol / ncs)
    extra["objective_evals_per_s"] = {"value": n_eval / dte, "ms_per_eval": dte * 1e3 / n_eval, "workload": "config2",
                                      "cpu_port_evals_per_s": 1.0 / (cpu_build + cpu_ll),
                                      "cpu_port_ms_per_eval": (cpu_build + cpu_ll) * 1e3, "cpu_cores": cores,
                                      "cpu_note": "NumPy restatement of the model build (oracle/ctmc_oracle.py, 1 thread; the reference's own "
                                                  "trans_emiss_calc takes 420-670 s, SURVEY 6) + C port of the forward sweep on all host threads"}
    extra["model_build_ms"] = {"one_set": float(np.median(build1)), "per_set_in_batch_of_1024": build1024 / 1024,
                               "batch_1024_ms": build1024, "cpu_port_ms": cpu_build * 1e3}
    extra["config5"] = {"sets": 1024, "columns": run2.ncol, "build_ms": build1024, "loglik_ms": sweep_s * 1e3,
                        "column_evaluations_per_s": 1024 * run2.ncol / (build1024 * 1e-3 + sweep_s),
                        "objective_evals_per_s": 1024 / (build1024 * 1e-3 + sweep_s),
                        "finite": bool(np.isfinite(ll1024).all())}
    extra["config3"] = run_config3(eng, barrier)
    return extra


def run_config3(eng, barrier):
    """Config 3: 100 Mb in 1 000 blocks at (n_int_AB, n_int_ABC) = (5, 5), K = 70 — forward
    log-likelihood and posterior (kept in HBM: 56 GB) on the lock-step FP64 tensor-core sweeps
    (csrc/lockstep.cu).  Ten copies of a seeded 10 Mb alignment sampled from the model itself."""
    from itrails_b200 import synth
    a, b, pi, _ = eng.build_model(synth.example_model_args(5)[None, :], 5, 5)
    a, b, pi = a[0], b[0], pi[0]
    K = a.shape[0]
    rng = np.random.default_rng(20261018 + 3)
    lens = synth.block_lengths(100, 10_000_000, rng)
    V = synth.alignment(a, b, pi, lens, 20261018 + 3)
    copies = 10
    lens_all = np.array([len(v) for v in V] * copies, dtype=np.int64)
    off = np.zeros(len(lens_all) + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens_all)
    eng.load_packed(np.tile(np.concatenate(V).astype(np.uint16), copies), off)
    n = int(off[-1])
    l0 = eng.lockstep_launch_count
    eng.loglik()
    eng.posterior(fetch=False)
    ll_ms, post_ms = [], []
    for _ in range(3):
        barrier()
        eng.loglik()
        ll_ms.append(eng.phase_ms("loglik"))
        eng.posterior(fetch=False)
        post_ms.append(eng.phase_ms("post_total"))
    ll, po = float(np.median(ll_ms)), float(np.median(post_ms))
    fl_ll, fl_po = n * (2.0 * K * K + 3 * K), n * (2.0 * (2 * K * K + 3 * K) + 3 * K)
    peak = measured_peaks()[1]
    return {"workload": "config3", "columns": n, "blocks": int(len(lens_all)), "K": int(K),
            "loglik_ms": ll, "posterior_ms": po, "posterior_hbm_gb": n * K * 8 / 1e9,
            "loglik_columns_per_s": n / (ll * 1e-3), "posterior_columns_per_s": n / (po * 1e-3),
            "loglik_tflops": fl_ll / (ll * 1e-3) / 1e12, "posterior_tflops": fl_po / (po * 1e-3) / 1e12,
            "posterior_frac_of_fp64_peak": fl_po / (po * 1e-3) / 1e12 / peak,
            "lockstep_launches": eng.lockstep_launch_count - l0,
            "note": "device time (CUDA events of the library's phases); flops per column as SURVEY 8(d)"}


if __name__ == "__main__":
    sys.exit(main())
