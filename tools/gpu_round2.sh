#!/usr/bin/env bash
# The round-2 record session on one B200 (through gpurun): parity tests, smoke, the driver's bench command + reference arm + ncu launch
# list, `ncu --set full` of one step's pruning launches, both Pupko designs, compute-sanitizer.  The multi-GPU lines come from
#   gpurun --gpus N -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
#                       bench.py --gpus N --steps 20 --warmup 5'
set -x
mkdir -p gpurun_out
timeout 1300 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -3 | tee gpurun_out/smoke.log
bash tools/gpu_final_n1.sh
bash tools/gpu_ncu_step.sh
bash tools/gpu_pupko2.sh
bash tools/gpu_sanitize.sh
timeout 1200 python tools/probes/pvalue_noise.py 2>&1 | tee gpurun_out/pvalue_noise.log
