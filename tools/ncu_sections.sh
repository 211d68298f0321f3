#!/bin/bash
# Which ncu sections can profile the Viterbi stream kernel?  (small run)
for sec in SpeedOfLight LaunchStats Occupancy MemoryWorkloadAnalysis ComputeWorkloadAnalysis SchedulerStats WarpStateStats InstructionStats SourceCounters PmSampling WorkloadDistribution; do
  ITR_VITERBI=stream timeout 120 ncu --section $sec --clock-control none --kernel-name regex:viterbi_stream_kernel --launch-count 1 \
      -f -o /tmp/sec_$sec python tools/vit_small.py 4 6000 > /tmp/sec_$sec.log 2>&1
  echo "$sec: $(grep -c LaunchFailed /tmp/sec_$sec.log) fail; $(grep -o '[0-9]* pass[es]*' /tmp/sec_$sec.log | head -1); $(grep -o 'viterbi small run.*' /tmp/sec_$sec.log)"
done
