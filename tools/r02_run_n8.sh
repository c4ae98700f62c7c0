#!/bin/bash
# round-2 evidence on 8 B200: the driver's own launch line for both arms, and the C++ single-process host on 8 GPUs
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/r02_topo8.txt 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > $O/r02_bench_n8.json 2> $O/r02_bench_n8.err; echo "bench n8 rc=$?"
tail -c 600 $O/r02_bench_n8.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 8 --steps 5 --warmup 3 > $O/r02_bench_ref_n8.json 2> $O/r02_bench_ref_n8.err; echo "ref n8 rc=$?"
{ timeout 300 ./gpucadforam_b200/gpucad_headless 4 512 8; timeout 300 ./gpucadforam_b200/gpucad_headless 5 1024 8; } > $O/r02_headless_8gpu.txt 2>&1
cat $O/r02_headless_8gpu.txt
tail -5 $O/r02_bench_n8.err
