#!/usr/bin/env bash
# ncu --set full of every pruning-kernel launch (factor tables + main pass) of ONE timed bench step at FAMILIES=${N:-1000000}
mkdir -p gpurun_out
N=${N:-1000000}
# prune_resident launches before the timed step: 6 evaluations of drop_failing_families + 1 warm-up step, LPS launches each
LPS=${LPS:-7}
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:prune_resident -s $((7 * LPS)) -c $LPS -f -o gpurun_out/prof_step_$N \
    python bench.py --families $N --steps 1 --warmup 1 --no-cpu-baseline --no-fit --no-weak > gpurun_out/ncu_step_$N.log 2>&1
tail -2 gpurun_out/ncu_step_$N.log | cut -c1-300
python tools/ncu_summary.py gpurun_out/prof_step_$N.ncu-rep gpurun_out/prof_step_$N.txt "ncu --set full, every pruning-kernel launch of one bench step, $N families on one B200"
grep -c "Kernel Name" gpurun_out/prof_step_$N.txt
