#!/bin/bash
# Round-end evidence: bench lines (no profiler), then the ncu launch list of the same
# command, then one `--set full` capture per dominant kernel (reports stay in /tmp on the
# box; only the CSV exports come back).  Run under gpurun:
#   gpurun --timeout 900 -- 'bash tools/capture_profiles.sh'
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err || exit 1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_launch.log 2>&1
# per step: 3 expm kernels, the Viterbi sweep, 16 checkpoint sweeps (8 length groups x 2 directions, shortest
# group first), 100 per-block pass-2 kernels, the log-likelihood sweep.  Skip the warm-up steps.
cap() {   # name, kernel regex, launches to skip
    timeout 240 ncu --set full --clock-control none --kernel-name "regex:$2" --launch-skip $3 --launch-count 1 \
        -f -o /tmp/prof_$1 $CMD > gpurun_out/ncu_full_$1.log 2>&1
    ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$1.csv 2>/dev/null
}
cap viterbi viterbi_stream_kernel 3
cap loglik forward_runs_kernel 3
cap expm 'expm_kernel<11' 3
cap sweep_fwd 'checkpoint_sweep_kernel<28, 0' 31     # 4th step, 8th launch = the longest group
cap sweep_bwd 'checkpoint_sweep_kernel<28, 1' 31
cap tiles posterior_tiles_kernel 399                # last block of the 4th step = a longest block
timeout 240 ncu --set full --import-source on --clock-control none --kernel-name regex:viterbi_stream_kernel --launch-skip 3 --launch-count 1 \
    -f -o /tmp/prof_vsrc $CMD > gpurun_out/ncu_full_vsrc.log 2>&1
ncu -i /tmp/prof_vsrc.ncu-rep --page source --csv > gpurun_out/ncu_source_viterbi.csv 2>/dev/null
ls -la gpurun_out | tail -20
