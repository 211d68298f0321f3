"""Does the host-side placement of the pinned ring matter on this box?  Device-to-host GB/s
into page-locked memory allocated (a) as the process finds itself, (b) after binding the
calling thread to the GPU's own CPUs (nvmlDeviceSetCpuAffinity), (c) bound to the OTHER CPUs."""
import os, subprocess, time
import torch
import pynvml

def rate(tag, nbytes=4 << 30, reps=3):
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    pin = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    best = 0.0
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        pin.copy_(dev, non_blocking=True); torch.cuda.synchronize()
        best = max(best, nbytes / (time.perf_counter() - t0) / 1e9)
    h2d = 0.0
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        dev.copy_(pin, non_blocking=True); torch.cuda.synchronize()
        h2d = max(h2d, nbytes / (time.perf_counter() - t0) / 1e9)
    print(f"{tag}: D2H {best:.1f} GB/s, H2D {h2d:.1f} GB/s, thread on cpus {sorted(os.sched_getaffinity(0))[:4]}..({len(os.sched_getaffinity(0))})", flush=True)
    del pin, dev

print(subprocess.run("nvidia-smi topo -m 2>&1 | head -14; lscpu | grep -i -E 'numa|socket|model name' ", shell=True, capture_output=True, text=True).stdout)
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
all_cpus = os.sched_getaffinity(0)
torch.cuda.init()
rate("as found")
try:
    pynvml.nvmlDeviceSetCpuAffinity(h)
    near = os.sched_getaffinity(0)
    rate("bound to the GPU's cpus")
    far = all_cpus - near
    if far:
        os.sched_setaffinity(0, far)
        rate("bound to the other cpus")
    os.sched_setaffinity(0, all_cpus)
    rate("unbound again")
except Exception as e:
    print("affinity:", repr(e))
