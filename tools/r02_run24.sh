#!/bin/bash
# fused primitive + retain: parity, config 2
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/tests24.log 2>&1; echo "tests rc=$?" >> $O/tests24.log; tail -12 $O/tests24.log
timeout 400 python tools/config_bench.py --configs 2 > $O/configs24.json 2> $O/configs24.err
python - <<'PY'
import json
for l in open("gpurun_out/configs24.json"):
    if l.startswith("{"):
        d=json.loads(l); print("  config",d["config"],{k:round(x["ms"],4) for k,x in d.items() if isinstance(x,dict) and "ms" in x}, d.get("fused_call",{}).get("same_mesh_as_legacy"), d.get("parity_full_size"))
PY
tail -3 $O/configs24.err
