#!/usr/bin/env bash
# Pupko session (round 2): parity tests touching reconstruction, then kernel times of both designs under the ncu launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "pupko or tables or small_trees or randomized or priors or multi_device or bucketed or larger_state or large_tree" 2>&1 | tail -8 | tee gpurun_out/pytest_pupko.log
for v in "CAFE_B200_PUPKO=1" "CAFE_B200_PUPKO=2" "CAFE_B200_PUPKO=2,CAFE_B200_TABLES=0"; do
  envs=$(echo "$v" | tr ',' ' ')
  tag=$(echo "$v" | tr -c 'A-Za-z0-9\n' '_')
  env $envs timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"pupko" --csv \
      --log-file gpurun_out/pupko2_ncu_$tag.csv python tools/gpu_pupko.py 2>&1 | tail -1
  python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/pupko2_ncu_$tag.csv")) if len(r)>10 and r[0].isdigit()]
# second reconstruct call only: take the last half of the launches
ids=sorted({int(r[0]) for r in rows}); half=ids[len(ids)//2:]
tot=collections.defaultdict(float)
for r in rows:
    if int(r[0]) in half:
        v=float(r[-1].replace(',','')); u=r[-2]
        if 'time' in r[-3]: tot['ms']+=v/1e6 if u in ('ns','nsecond') else v*{'us':1e-3,'usecond':1e-3,'ms':1,'msecond':1}.get(u,1)
        if 'bytes_read' in r[-3]: tot['GB_read']+=v*{'byte':1e-9,'Kbyte':1e-6,'Mbyte':1e-3,'Gbyte':1}.get(u,1e-9)
        if 'bytes_write' in r[-3]: tot['GB_write']+=v*{'byte':1e-9,'Kbyte':1e-6,'Mbyte':1e-3,'Gbyte':1}.get(u,1e-9)
print("== $v: %d launches per reconstruction, %.1f ms, DRAM read %.1f GB, write %.1f GB" % (len(half), tot['ms'], tot['GB_read'], tot['GB_write']))
PY
done
