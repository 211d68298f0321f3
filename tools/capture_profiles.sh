#!/bin/bash
# Round-end evidence: bench lines (no profiler), then the ncu launch list of the same
# command, then one `--set full` capture of the five dominant kernels.  Run under gpurun:
#   gpurun --timeout 1500 -- 'bash tools/capture_profiles.sh'
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err || exit 1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on \
    --kernel-name regex:'viterbi_spec_kernel|checkpoint_sweep_kernel|posterior_tiles_kernel|forward_runs_kernel|expm_kernel' \
    --launch-skip 12 --launch-count 14 -f -o gpurun_out/prof_final $CMD > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/prof_final.ncu-rep --page raw --csv > gpurun_out/ncu_final_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_final.ncu-rep --page source --csv --kernel-name regex:viterbi_spec_kernel > gpurun_out/ncu_final_source_viterbi.csv 2>/dev/null
ls -la gpurun_out | tail -12
