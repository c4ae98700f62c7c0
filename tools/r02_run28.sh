#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 5 --warmup 3 > $O/bench28_n$N.json 2> $O/bench28_n$N.err; echo "bench n$N rc=$?"
tail -4 $O/bench28_n$N.err
python - <<PY
import json
L=[l for l in open("gpurun_out/bench28_n$N.json") if l.startswith("{")]
d=json.loads(L[-1])
print({k:d[k] for k in ("n_gpus","ms_per_step","value")})
e=d["e2e"]; print("e2e ms", e["ms_per_step"], "blocking", e["blocking_call"]["ms_per_step"], e["mode"][:60])
print("probe", e.get("h2d_probe"))
for k in ("fast_field","config4"):
    x=d.get(k) or {}
    print(k, x.get("ms_per_step"), (x.get("e2e") or {}).get("ms_per_step"), (x.get("e2e") or {}).get("blocking_call_ms"))
PY
