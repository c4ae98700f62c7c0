#!/bin/bash
# round-2 evidence on one B200: parity suite, both bench arms, configs 1/2/5 with full-size parity, launch lists, ncu captures
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/r02_gpu_tests.log 2>&1; echo "tests rc=$?" >> $O/r02_gpu_tests.log
tail -5 $O/r02_gpu_tests.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/r02_bench_ours.json 2> $O/r02_bench_ours.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/r02_bench_reference.json 2> $O/r02_bench_ref.err; echo "ref rc=$?"
timeout 300 python tools/config_bench.py > $O/r02_configs.json 2> $O/r02_configs.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r02_launches.csv python bench.py --profile --steps 2 --warmup 3 > $O/r02_launches.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r02_launches_fast.csv python bench.py --profile --fast-field --steps 2 --warmup 3 > $O/r02_launches_fast.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:svl_field_fast -s 2 -c 1 -f -o $O/r02_fast_field_v3 python bench.py --profile --fast-field --steps 1 --warmup 3 > $O/ncu8a.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:svl_field_tile -s 2 -c 1 -f -o $O/r02_exact_field python bench.py --profile --steps 1 --warmup 3 > $O/ncu8b.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mc_fused -s 3 -c 1 -f -o $O/r02_mc_fused_final python bench.py --profile --steps 1 --warmup 3 > $O/ncu8c.log 2>&1
./gpucadforam_b200/gpucad_headless 3 256 --full > $O/r02_headless.txt 2>&1
./gpucadforam_b200/gpucad_headless 4 256 4 >> $O/r02_headless.txt 2>&1
./gpucadforam_b200/gpucad_headless 5 512 4 >> $O/r02_headless.txt 2>&1
bash tools/sanitize.sh > /dev/null 2>&1; tail -20 $O/compute_sanitizer.txt
