#!/usr/bin/env bash
# L2 experiments on the 1 M-family bench step (quick bench lines, no baseline / fits / weak field):
#   base | CAFE_B200_L2_WINDOW=slots | =arena | library built with -DCAFE_TABLE_STORE_CS=1 (cafe5_b200/libcafe_b200_cs.so)
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-fit --no-weak > gpurun_out/l2_$name.json 2> gpurun_out/l2_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/l2_%s.json" % name).read().strip().splitlines()[-1])
    print("%-8s ms_per_step %.3f  e2e_ms %.3f  kernel_ms %.3f" % (name, d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["ms_per_step_kernel"]))
except Exception as e:
    print(name, "FAILED", e)
PY
}
run base X=1
run slots CAFE_B200_L2_WINDOW=slots
run arena CAFE_B200_L2_WINDOW=arena
cp cafe5_b200/libcafe_b200.so /tmp/libcafe_b200_default.so
cp cafe5_b200/libcafe_b200_cs.so cafe5_b200/libcafe_b200.so
run cs X=1
cp /tmp/libcafe_b200_default.so cafe5_b200/libcafe_b200.so
