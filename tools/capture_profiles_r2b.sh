#!/bin/bash
# Round-2 (second session) evidence: run under gpurun from the repo root; everything lands in gpurun_out/.
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/r2d_tests.log
python bench.py > $O/bench_r2_final_n1.json 2> $O/bench_r2_final_n1.err
python bench.py --impl reference > $O/bench_r2_final_reference_arm.json 2> /dev/null
{ for d in "5 5" "3 4" "3 5" "5 6"; do python tools/time_config3.py $d | grep -v "norot=1" | grep -v "LOCKSTEP=1 "; done; } > $O/config3_r2.txt 2>&1
python tools/time_lockstep.py 20000 592 1184 2368 > $O/lockstep_steps_r2.txt 2>&1
python tools/time_vit_modes.py 0.2 > $O/vit_modes_r2.txt 2>&1
python tools/time_share_step.py 1.0 - 8,4,2,1 > $O/share_step_r2.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:lockstep --launch-skip 3 --launch-count 2 -o $O/ncu_lockstep_r2 -f python tools/time_lockstep.py 20000 592 > $O/ncu_lockstep.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench_r2_final.csv python bench.py --steps 2 --warmup 1 --no-extra > $O/ncu_bench.log 2>&1
tail -2 $O/r2d_tests.log
