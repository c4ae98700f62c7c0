#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/topo2.txt 2>&1
{ ./gpucadforam_b200/gpucad_headless 4 256 2; ./gpucadforam_b200/gpucad_headless 5 512 2; ./gpucadforam_b200/gpucad_headless 4 512 2; ./gpucadforam_b200/gpucad_headless 5 768 2; } > $O/headless_2gpu.txt 2>&1
cat $O/headless_2gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err; echo "ref n2 rc=$?"
timeout 300 python tools/config_bench.py --configs 2 > $O/configs5.json 2> $O/configs5.err
tail -3 $O/bench_n2.err
