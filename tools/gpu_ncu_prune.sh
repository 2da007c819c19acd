#!/usr/bin/env bash
# ncu --set full capture of the pruning kernel in the bench configuration (skips the 400 single-matrix launches of the workload generator)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:prune_ -s 1 -c 1 -f -o gpurun_out/prof_prune \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fit > gpurun_out/ncu_prune.log 2>&1
tail -2 gpurun_out/ncu_prune.log | cut -c1-300
