#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/tests15.log 2>&1; echo "tests rc=$?" >> $O/tests15.log
tail -6 $O/tests15.log
timeout 300 python tools/config_bench.py --configs 2 > $O/configs15.json 2> $O/configs15.err
GCB_RETAIN_SCALAR=1 timeout 300 python tools/config_bench.py --configs 2 > $O/configs15_scalar.json 2>> $O/configs15.err
python - <<'PY'
import json
for f in ("gpurun_out/configs15.json","gpurun_out/configs15_scalar.json"):
    for l in open(f):
        d=json.loads(l); print(f, d['legacy_calls']['ms'], d['enqueue_only_calls']['ms'], d['reference_kernels']['ms'], d['parity_full_size'])
PY
./gpucadforam_b200/gpucad_headless 4 128 3 | tail -1; ./gpucadforam_b200/gpucad_headless 5 128 3 | tail -1
