#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
for tc in 1024 2048 4096 8192; do
  GCB_MC_TILE_CELLS=$tc timeout 300 python tools/config_bench.py --configs 1,2 --steps 10 > $O/configs25_$tc.json 2> $O/configs25.err
  python - <<PY
import json
for l in open("gpurun_out/configs25_$tc.json"):
    if l.startswith("{"):
        d=json.loads(l); print("tile_cells $tc config",d["config"],{k:round(x["ms"],4) for k,x in d.items() if isinstance(x,dict) and "ms" in x})
PY
done
