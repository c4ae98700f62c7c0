#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/tests3.log 2>&1; echo "tests rc=$?" >> $O/tests3.log
tail -30 $O/tests3.log
timeout 300 python tools/config_bench.py > $O/configs3.json 2> $O/configs3.err; echo "configs rc=$?"
./gpucadforam_b200/gpucad_headless 4 256 4 > $O/headless4.txt 2>&1; ./gpucadforam_b200/gpucad_headless 5 512 4 >> $O/headless4.txt 2>&1; cat $O/headless4.txt
