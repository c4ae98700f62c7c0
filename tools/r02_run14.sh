#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" > $O/r02_gpu_tests_full.log; tail -4 $O/r02_gpu_tests_full.log; grep "512^3 fast field" $O/r02_gpu_tests_full.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/r02_bench_ours.json 2> $O/r02_bench_ours.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/r02_bench_reference.json 2> $O/r02_bench_ref.err; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
