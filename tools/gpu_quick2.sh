#!/usr/bin/env bash
# quick GPU iteration (round 2): parity tests + short bench of each variant named in $VARIANTS ("ENV=VAL,ENV=VAL;..."), FAMILIES=${N:-1000000}
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != "1" ]; then
timeout 900 python -m pytest tests -m gpu -x -q ${TESTS_K:+-k "$TESTS_K"} 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
fi
IFS=';' read -ra VS <<< "${VARIANTS:-default}"
for v in "${VS[@]}"; do
  envs=$(echo "$v" | tr ',' ' ')
  [ "$v" = "default" ] && envs=""
  for n in ${SIZES:-1000000 125000}; do
  echo "== variant: $v families $n"
  env $envs timeout 600 python bench.py --families $n --steps 6 --warmup 3 --no-cpu-baseline --no-fit --no-weak 2> gpurun_out/bench.err | tee "gpurun_out/bench_quick_$(echo "$v" | tr -c 'A-Za-z0-9\n' '_')_$n.json" | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=d['roofline']
print('value %.0f e2e %.0f ms/step %.2f prune ms %.2f TF %.2f frac %.3f (equiv %.1f TF) mat ms %.2f launches/step %d'%(d['value'],d['e2e']['value'],d['ms_per_step'],r['ms_per_step_kernel'],r['achieved'],r['frac'],r['pattern_reuse']['equivalent_TFLOPs'],r['matrix_gen']['ms_per_launch'],r['launches_per_step']), d['clocks'], d['result']['neg_lnl'])
"
  tail -3 gpurun_out/bench.err
  done
done
