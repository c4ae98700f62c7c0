#!/bin/bash
# round-2 GPU session 2: parity suite (new tests), bench with pipelined e2e, fast-field kernel v2 profile
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -s > $O/tests2.log 2>&1; echo "tests rc=$?" >> $O/tests2.log
tail -15 $O/tests2.log
timeout 600 python bench.py > $O/bench2_ours.json 2> $O/bench2_ours.err; echo "bench rc=$?"
timeout 120 python bench.py --profile --fast-field > $O/fast2_profile.json 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:svl_field_fast -s 2 -c 1 -f -o $O/r02_fast_field_v2 python bench.py --profile --fast-field --steps 1 --warmup 3 > $O/ncu_fast2.log 2>&1
ls -la $O | tail -8
