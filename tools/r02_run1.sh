#!/bin/bash
# round-2 GPU session 1: parity suite, bench (both arms), launch list, ncu captures of the fast field kernel and the extraction kernel
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $O/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -s > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
tail -5 $O/tests.log
timeout 600 python bench.py > $O/bench_ours.json 2> $O/bench_ours.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
GCB_SVL_FAST_MINB=4 timeout 120 python bench.py --profile --fast-field > $O/fast_minb4.json 2>&1
timeout 120 python bench.py --profile --fast-field > $O/fast_minb3.json 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_fast.csv python bench.py --profile --fast-field --steps 2 --warmup 3 > $O/launches_fast.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:svl_field_fast -s 2 -c 1 -f -o $O/r02_fast_field python bench.py --profile --fast-field --steps 1 --warmup 3 > $O/ncu_fast.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mc_fused -s 3 -c 1 -f -o $O/r02_mc_fused python bench.py --profile --steps 1 --warmup 3 > $O/ncu_mc.log 2>&1
ls -la $O
