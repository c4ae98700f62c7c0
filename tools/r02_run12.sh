#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke12.txt 2>&1; tail -2 $O/smoke12.txt
for tc in 2048 4096 8192 16384; do
  echo "tile_cells=$tc"; GCB_MC_TILE_CELLS=$tc timeout 200 python tools/config_bench.py --configs 2,5 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'], round(d['legacy_calls']['ms'],4), round(d.get('fused_call',{}).get('ms',0),4), round(d.get('enqueue_only_calls',{}).get('ms',0),4))"
done > $O/tile_sweep12.txt 2>&1
cat $O/tile_sweep12.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gcb -c 40 --csv --log-file $O/r02_launches.csv python bench.py --profile --steps 2 --warmup 3 > $O/l12a.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gcb -c 40 --csv --log-file $O/r02_launches_fast.csv python bench.py --profile --fast-field --steps 2 --warmup 3 > $O/l12b.log 2>&1
wc -l $O/r02_launches.csv $O/r02_launches_fast.csv
