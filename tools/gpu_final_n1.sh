#!/usr/bin/env bash
# round-2 record run on one GPU: the driver's bench command, the reference arm, the ncu launch list of the same command
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv
timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2> gpurun_out/r02_bench_reference.err | tee gpurun_out/r02_bench_reference_n1.json | cut -c1-400
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/r02_bench_n1.err | tee gpurun_out/r02_bench_n1.json | cut -c1-600
tail -3 gpurun_out/r02_bench_n1.err
SIZES=1000000 bash tools/gpu_launchlist.sh
cp gpurun_out/launches_1000000.csv gpurun_out/r02_launches_n1.csv
