#!/usr/bin/env bash
# One GPU session for the matrix-kernel and Pupko-geometry changes: parity tests, quick bench of both matrix kernels, Pupko kernel
# time (ncu launch list, one metric) of both geometries with a checksum of the reconstructed states.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
for v in "" "CAFE_B200_MATGEN=entry" "CAFE_B200_MATGEN=rows"; do
  echo "== bench ${v:-default}"
  env $v timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-fit 2> gpurun_out/bench.err | tee "gpurun_out/bench_quick_matgen_${v##*=}.json" | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f e2e %.0f ms/step %.2f prune ms %.2f TF %.2f frac %.3f mat ms %.3f launches %d'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['ms_per_step_kernel'],d['roofline']['achieved'],d['roofline']['frac'],d['roofline']['matrix_gen']['ms_per_launch'],d['gpu_launches']), d['clocks'], d['result'])
"
  tail -3 gpurun_out/bench.err
done
for t in 512 256; do
  CAFE_B200_PUPKO_THREADS=$t timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,launch__registers_per_thread \
      --clock-control none -k regex:pupko --csv --log-file gpurun_out/pupko_ncu_$t.csv python tools/gpu_pupko.py 2>&1 | tail -2
  grep -o '"gpu__time_duration.sum","ns","[0-9]*"\|fp64_cycles_active[^,]*,"%","[0-9.]*"\|issue_active[^,]*,"%","[0-9.]*"\|pipe_alu[^,]*,"%","[0-9.]*"' gpurun_out/pupko_ncu_$t.csv | tail -4
done
