#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 5 --warmup 3 > $O/r02_bench_n4.json 2> $O/r02_bench_n4.err; echo "bench n4 rc=$?"
tail -c 900 $O/r02_bench_n4.json; echo
tail -5 $O/r02_bench_n4.err
nvidia-smi --query-gpu=memory.used --format=csv | head -5
