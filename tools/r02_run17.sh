#!/bin/bash
# vectorised primitive kernels: parity + config 2 A/B
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/tests17.log 2>&1; echo "tests rc=$?" >> $O/tests17.log
tail -25 $O/tests17.log

timeout 300 python tools/config_bench.py --configs 2 > $O/configs17.json 2> $O/configs17.err
GCB_PRIM_SCALAR=1 timeout 300 python tools/config_bench.py --configs 2 > $O/configs17_scalar.json 2>> $O/configs17.err
python - <<'PY'
import json
for f in ("gpurun_out/configs17.json","gpurun_out/configs17_scalar.json"):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); print(f, d['legacy_calls']['ms'], d['enqueue_only_calls']['ms'], d['reference_kernels']['ms'], d['parity_full_size'])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches17.csv python tools/config_bench.py --configs 2 --steps 2 --warmup 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches17.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
for r in rows[1:40]: print(r[ki][:70], r[vi])
PY
