#!/bin/bash
# separable TPMS kernel: parity, full tests, config 1 with / without the tables
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== per-point kernel"; GCB_TPMS_NO_TABLES=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "create_lattice" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -q > $O/tests19.log 2>&1; echo "tests rc=$?" >> $O/tests19.log; tail -6 $O/tests19.log
timeout 300 python tools/config_bench.py --configs 1 > $O/configs19.json 2> $O/configs19.err
GCB_TPMS_NO_TABLES=1 timeout 300 python tools/config_bench.py --configs 1 > $O/configs19_pp.json 2>> $O/configs19.err
python - <<'PY'
import json
for f in ("gpurun_out/configs19.json","gpurun_out/configs19_pp.json"):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); print(f, d['legacy_calls']['ms'], d['fused_call']['ms'], d['reference_kernels']['ms'], d['parity_full_size'])
PY
