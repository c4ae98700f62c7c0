#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/tests6.log 2>&1; echo "tests rc=$?" >> $O/tests6.log
tail -12 $O/tests6.log
timeout 120 python bench.py --profile --fast-field > $O/mc6_new.json 2>&1
GCB_MC_NO_RAW_STAGE=1 timeout 120 python bench.py --profile --fast-field > $O/mc6_old.json 2>&1
cat $O/mc6_new.json $O/mc6_old.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mc_fused -s 3 -c 1 -f -o $O/r02_mc_fused_v2 python bench.py --profile --fast-field --steps 1 --warmup 3 > $O/ncu_mc6.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_cfg2.csv python tools/config_bench.py --configs 2 --steps 1 --warmup 1 > $O/launches_cfg2.log 2>&1
