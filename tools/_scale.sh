#!/bin/bash
# weak scaling N=1,2,4,8 of the default workload + BASELINE config 4 (2048^3 over 8 and 4 GPUs); run under gpurun --gpus 8
set -u
mkdir -p gpurun_out
run() { # n extra...
  n=$1; shift
  if [ "$n" = 1 ]; then timeout 400 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline "$@"
  else timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 "$@"; fi
}
for n in 1 2 4 8; do run $n 2>gpurun_out/scale_n$n.err | tail -1 > gpurun_out/scale_weak_n$n.json; echo "weak n=$n: $(cut -c1-200 gpurun_out/scale_weak_n$n.json)"; done
run 8 --fine 2048 --strong 2>gpurun_out/strong_n8.err | tail -1 > gpurun_out/strong_2048_n8.json; echo "strong 2048 n=8: $(cut -c1-260 gpurun_out/strong_2048_n8.json)"
run 4 --fine 2048 --strong 2>gpurun_out/strong_n4.err | tail -1 > gpurun_out/strong_2048_n4.json; echo "strong 2048 n=4: $(cut -c1-260 gpurun_out/strong_2048_n4.json)"
