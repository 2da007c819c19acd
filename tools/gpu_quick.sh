#!/usr/bin/env bash
# quick GPU iteration: parity tests + short bench of each variant named in $VARIANTS ("ENV=VAL,ENV=VAL;...")
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
IFS=';' read -ra VS <<< "${VARIANTS:-default}"
for v in "${VS[@]}"; do
  envs=$(echo "$v" | tr ',' ' ')
  [ "$v" = "default" ] && envs=""
  echo "== variant: $v"
  env $envs timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-fit 2> gpurun_out/bench.err | tee "gpurun_out/bench_quick_$(echo "$v" | tr -c 'A-Za-z0-9\n' '_').json" | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f e2e %.0f ms/step %.2f prune ms %.2f TF %.2f frac %.3f mat ms %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['ms_per_step_kernel'],d['roofline']['achieved'],d['roofline']['frac'],d['roofline']['matrix_gen']['ms_per_launch']), d['clocks'], d['result'])
"
  tail -3 gpurun_out/bench.err
done
