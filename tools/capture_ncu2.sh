#!/bin/bash
# Second part of the round-end evidence (see capture_profiles.sh): `--set full` of the Viterbi
# sweep fails with LaunchFailed on the bench command (every section works on its own), so
# the sections are requested explicitly; sweeps and expm are selected by launch index.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
SECS="--section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section SchedulerStats --section WarpStateStats --section LaunchStats --section Occupancy --section InstructionStats"
timeout 300 ncu $SECS --clock-control none --kernel-name regex:viterbi_stream_kernel --launch-skip 3 --launch-count 1 \
    -f -o /tmp/prof_viterbi $CMD > gpurun_out/ncu_full_viterbi.log 2>&1
ncu -i /tmp/prof_viterbi.ncu-rep --page raw --csv > gpurun_out/ncu_raw_viterbi.csv 2>/dev/null
cap() {   # name, kernel regex, launches to skip, count
    timeout 240 ncu --set full --clock-control none --kernel-name "regex:$2" --launch-skip $3 --launch-count $4 \
        -f -o /tmp/prof_$1 $CMD > gpurun_out/ncu_full_$1.log 2>&1
    ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$1.csv 2>/dev/null
}
cap sweeps checkpoint_sweep_kernel 62 2      # 4th step: backward and forward sweep of the longest group
cap expm expm_kernel 9 3
# source-level view of the Viterbi sweep on a stand-alone run of the same shape
ITR_VITERBI=stream timeout 300 ncu --section SourceCounters --section WarpStateStats --import-source on --clock-control none \
    --kernel-name regex:viterbi_stream_kernel --launch-skip 2 --launch-count 1 -f -o /tmp/prof_vsrc python tools/time_vit.py 100 100000 > gpurun_out/ncu_full_vsrc.log 2>&1
ncu -i /tmp/prof_vsrc.ncu-rep --page source --csv > gpurun_out/ncu_source_viterbi.csv 2>/dev/null
ls -la gpurun_out | tail; tail -4 gpurun_out/ncu_full_viterbi.log
