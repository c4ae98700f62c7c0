#!/bin/bash
# final single-GPU evidence refresh: GPU tests (incl. the 2048-wide config 4 slab), configs 1/2/5, both bench arms
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > $O/tests16.log 2>&1; echo "tests rc=$?" >> $O/tests16.log
tail -14 $O/tests16.log
timeout 400 python tools/config_bench.py > $O/configs16.json 2> $O/configs16.err; echo "configs rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench16_ref.json 2> $O/bench16_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench16.json 2> $O/bench16.err; echo "ours rc=$?"
timeout 600 python bench.py --fast-field --no-extra --steps 20 --warmup 3 > $O/bench16_fast.json 2> $O/bench16_fast.err; echo "fast rc=$?"
python - <<'PY'
import json
def last(f):
    L=[l for l in open(f) if l.startswith("{")]
    return json.loads(L[-1]) if L else None
for f in ("gpurun_out/bench16_ref.json","gpurun_out/bench16.json","gpurun_out/bench16_fast.json"):
    d=last(f)
    if d: print(f, d.get("ms_per_step"), d.get("value"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"))
PY
