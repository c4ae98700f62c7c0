#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/tests4.log 2>&1; echo "tests rc=$?" >> $O/tests4.log
tail -30 $O/tests4.log
timeout 300 python tools/config_bench.py --configs 2 > $O/configs4.json 2> $O/configs4.err; echo "configs rc=$?"
timeout 300 python tools/config_bench.py --configs 2 --async-fields > $O/configs4_async.json 2>> $O/configs4.err
./gpucadforam_b200/gpucad_headless 3 256 --full > $O/headless_full.txt 2>&1; cat $O/headless_full.txt
