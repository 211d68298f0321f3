#!/bin/bash
# Round-2 evidence, one GPU box: tests, the bench line of record, its ncu launch list, `--set
# full` captures of the dominant kernels at config-4 launch size, the CPU arms.
# Usage (from the repo root): gpurun --timeout 1700 -- 'bash tools/capture_profiles_r2.sh'
O=gpurun_out/r2p
mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log)
tail -3 $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc $?"
timeout 600 python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; echo "reference arm rc $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > $O/launches_bench.log 2>&1
cap() {   # name, kernel regex, what, launches to skip, count
    timeout 400 ncu --set full --import-source on --clock-control none -k "regex:$2" -s $4 -c $5 -f -o $O/ncu_$1 \
        python tools/prof_config4.py $3 1.0 1 > $O/ncu_$1.log 2>&1
    ncu -i $O/ncu_$1.ncu-rep --page raw --csv > $O/ncu_raw_$1.csv 2>/dev/null
    ncu -i $O/ncu_$1.ncu-rep --page source --csv > $O/ncu_source_$1.csv 2>/dev/null
    rm -f $O/ncu_$1.ncu-rep
}
cap tiles_mma posterior_tiles_mma posterior 0 1
cap sweeps checkpoint_sweep posterior 0 2
cap vcheck viterbi_check viterbi 0 1
cap loglik forward_runs loglik 0 1
python tools/ncu_summary.py $O/ncu_raw_tiles_mma.csv $O/ncu_raw_sweeps.csv $O/ncu_raw_vcheck.csv $O/ncu_raw_loglik.csv | tee $O/ncu_summary.txt
timeout 600 python tools/time_reference_python.py 100000 > $O/reference_python.json 2> $O/reference_python.err; tail -30 $O/reference_python.json
