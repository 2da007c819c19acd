#!/usr/bin/env bash
# One GPU session: parity tests, smoke, bench, ncu launch list and full captures of the two hot kernels.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv
timeout 120 python tools/gpu_peaks.py 2>&1 | tee gpurun_out/peaks.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 600 python bench.py --steps 8 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
CAFE_B200_PRUNE=dfma timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fit 2>/dev/null | tee gpurun_out/bench_dfma.json
if [ "${SKIP_NCU:-0}" != "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 836 -c 48 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fit > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:prune_ -s 1 -c 1 -f -o gpurun_out/prof_prune \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fit > gpurun_out/ncu_prune.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:matrix_gen -s 407 -c 1 -f -o gpurun_out/prof_matrix \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fit > gpurun_out/ncu_matrix.log 2>&1
fi
ls -la gpurun_out
# every kernel variant through the parity tests (the default run above exercises only the geometry the host picks)
for v in "CAFE_B200_PRUNE=dfma" "CAFE_B200_PRUNE=stream" "CAFE_B200_RESIDENT_WN=4" "CAFE_B200_PUPKO_THREADS=256" "CAFE_B200_MATGEN=entry"; do
  echo "== $v"; env $v timeout 400 python -m pytest tests -m gpu -x -q -k "config1 or config2 or small_trees or cliff or config3 or randomized or config5 or pupko" 2>&1 | tail -2 | tee -a gpurun_out/pytest_gpu_variants.log
done
