"""One line per captured launch from `ncu --page raw --csv` exports (profiles/ncu_raw_*_r1.csv):
duration, DRAM bytes, registers, grid, FP64 pipe, issue slots, active warps, shared-memory
wavefronts, L1 hit rate.  Usage: python tools/ncu_summary.py profiles/ncu_raw_*_r1.csv"""
import csv, re, sys

M = {"t": "gpu__time_duration.sum", "r": "dram__bytes_read.sum", "w": "dram__bytes_write.sum",
     "regs": "launch__registers_per_thread", "grid": "launch__grid_size", "block": "launch__block_size",
     "fp64": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
     "issue": "sm__inst_issued.avg.pct_of_peak_sustained_active",
     "warps": "sm__warps_active.avg.pct_of_peak_sustained_active",
     "smem": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
     "l1": "l1tex__t_sector_hit_rate.pct"}
SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def num(d, u, k):
    if M[k] not in d or d[M[k]] in ("", "n/a"):
        return float("nan")
    return float(d[M[k]].replace(",", "")) * SCALE.get(u.get(M[k], ""), 1.0)


for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], dict(zip(rows[0], rows[1]))
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "").replace("itr::", "").replace("(int)", "")
        print(f"{name:<42s} {num(d, units, 't'):8.3f} ms  dram {num(d, units, 'r') + num(d, units, 'w'):8.2f} MB  "
              f"regs {int(num(d, units, 'regs')):4d}  grid {int(num(d, units, 'grid')):5d} x {int(num(d, units, 'block')):4d}  "
              f"fp64 pipe {num(d, units, 'fp64'):6.2f}%  issue {num(d, units, 'issue'):6.2f}%  warps active {num(d, units, 'warps'):6.2f}%  "
              f"smem wavefronts {num(d, units, 'smem'):6.2f}%  L1 hit {num(d, units, 'l1'):6.2f}%")
