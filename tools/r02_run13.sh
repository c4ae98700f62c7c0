#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/tests13.log 2>&1; echo "tests rc=$?" >> $O/tests13.log
tail -4 $O/tests13.log
for i in 1 2 3; do timeout 120 python bench.py --profile --fast-field 2>&1 | tail -1; done > $O/mc13.json
cat $O/mc13.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"mc_fused|minmax|svl_field" -c 40 --csv --log-file $O/r02_launches.csv python bench.py --profile --steps 2 --warmup 3 > $O/l13a.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"mc_fused|minmax|svl_field" -c 40 --csv --log-file $O/r02_launches_fast.csv python bench.py --profile --fast-field --steps 2 --warmup 3 > $O/l13b.log 2>&1
wc -l $O/r02_launches.csv $O/r02_launches_fast.csv
