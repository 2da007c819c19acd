#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:prune_ -s 1 -c 1 -f -o gpurun_out/prof_prune \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_prune.log 2>&1
tail -2 gpurun_out/ncu_prune.log | cut -c1-300
