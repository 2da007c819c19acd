#!/usr/bin/env bash
# Pupko kernel session: full parity suite, kernel time at the bench shard size (ncu launch list), one --set full capture.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
for t in 512 256; do
  CAFE_B200_PUPKO_THREADS=$t timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,launch__registers_per_thread \
      --clock-control none -k regex:pupko --csv --log-file gpurun_out/pupko_ncu_$t.csv python tools/gpu_pupko.py 2>&1 | tail -2
  grep -o '"gpu__time_duration.sum","ns","[0-9]*"\|fp64_cycles_active[^,]*,"%","[0-9.]*"\|issue_active[^,]*,"%","[0-9.]*"\|pipe_alu[^,]*,"%","[0-9.]*"' gpurun_out/pupko_ncu_$t.csv | tail -4
done
if [ "${SKIP_FULL:-0}" != "1" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pupko -s 1 -c 1 -f -o gpurun_out/prof_pupko python tools/gpu_pupko.py > gpurun_out/ncu_pupko.log 2>&1
tail -2 gpurun_out/ncu_pupko.log | cut -c1-200
fi
