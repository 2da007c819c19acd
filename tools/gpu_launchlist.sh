#!/usr/bin/env bash
# ncu launch list (gpu__time_duration only) of the pruning-kernel launches of bench steps, for FAMILIES in $SIZES
mkdir -p gpurun_out
for n in ${SIZES:-1000000 125000}; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"prune_resident|matrix_gen|finish|final_sum|pow_table" --csv \
      --log-file gpurun_out/launches_$n.csv python bench.py --families $n --steps 2 --warmup 1 --no-cpu-baseline --no-fit --no-weak > gpurun_out/bench_under_ncu_$n.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_$n.csv")) if len(r)>10 and r[0].isdigit()]
# columns: ID, Process ID, Process Name, Host Name, Kernel Name, Context, Stream, Block Size, Grid Size, Device, CC, Section Name, Metric Name, Metric Unit, Metric Value
out=[(r[4].split('(')[0][-60:], r[8], float(r[-1].replace(',',''))/1e6) for r in rows]
# the last step: walk back from the end to the previous final_sum
last=[]
seen=0
for name,grid,ms in reversed(out):
    if 'final_sum' in name:
        seen+=1
        if seen==2: break
    last.append((name,grid,ms))
print("== families $n: launches of the last step")
for name,grid,ms in reversed(last): print("%-62s grid %-14s %9.3f ms"%(name,grid,ms))
PY
done
