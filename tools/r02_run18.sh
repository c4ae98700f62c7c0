#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for c in 1 5; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches18_c$c.csv python tools/config_bench.py --configs $c --steps 2 --warmup 1 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches18_c$c.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
print("config $c")
for r in rows[1:]:
    if "at::" in r[ki]: continue
    print("  ", r[ki][:90], r[vi])
PY
done
