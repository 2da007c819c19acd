#!/usr/bin/env bash
# subtree-pattern tables: parity tests, then the bench with and without them
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "tables or config5 or multi_device or randomized or small_trees" 2>&1 | tail -15 | tee gpurun_out/pytest_tables.log
for v in "CAFE_B200_TABLES=0" "CAFE_B200_TABLE_FRAC=0.25" "CAFE_B200_TABLE_FRAC=0.5" "CAFE_B200_TABLE_FRAC=0.75" "CAFE_B200_TABLE_FRAC=0.9"; do
  echo "== $v"
  env $v timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-fit 2> gpurun_out/bench.err | tee "gpurun_out/bench_tables_${v##*=}.json" | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f e2e %.0f ms/step %.2f prune ms %.2f mat ms %.3f launches %d'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['ms_per_step_kernel'],d['roofline']['matrix_gen']['ms_per_launch'],d['gpu_launches']), d['result'])
"
  tail -3 gpurun_out/bench.err
done
