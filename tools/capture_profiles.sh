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
# per step the bench launches: 3 expm kernels (model build), the Viterbi sweep, 16 checkpoint sweeps (8 length
# groups x 2 directions), 100 per-block pass-2 kernels, the log-likelihood sweep.  Skip the 3 warm-up steps.
ncu --set full --clock-control none --import-source on \
    --kernel-name regex:'viterbi_s.*_kernel|forward_runs_kernel|expm_kernel' \
    --launch-skip 15 --launch-count 5 -f -o gpurun_out/prof_final_a $CMD > gpurun_out/ncu_full_a.log 2>&1
ncu --set full --clock-control none --import-source on \
    --kernel-name regex:'checkpoint_sweep_kernel|posterior_tiles_kernel' \
    --launch-skip 348 --launch-count 116 -f -o gpurun_out/prof_final_b $CMD > gpurun_out/ncu_full_b.log 2>&1
ls -la gpurun_out | tail -12
