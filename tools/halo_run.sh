#!/bin/bash
# Stored-field z-slabs with the +z halo layer over NCCL point-to-point (BASELINE config 5 sharded): GPU tests, then the tool at
# N = 1 and N = $2 with the byte-for-byte check against the single-GPU mesh.   gpurun --gpus 2 -- 'bash tools/halo_run.sh t22 2'
# third argument "only": just the N = $2 run (a call on 8 GPUs is charged 8x).
t=${1:-halo}; n=${2:-2}; only=${3:-}
mkdir -p gpurun_out
if [ -z "$only" ]; then
  (timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/${t}_tests.log
  tail -2 gpurun_out/${t}_tests.log
  timeout 300 python tools/config5_multi.py --check 2>gpurun_out/${t}_config5_n1.err | tail -1 > gpurun_out/${t}_config5_n1.json
  cat gpurun_out/${t}_config5_n1.json
fi
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,P2P timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
  --master-port 29517 tools/config5_multi.py --check > gpurun_out/${t}_config5_n$n.out 2>gpurun_out/${t}_config5_n$n.err
grep '^{' gpurun_out/${t}_config5_n$n.out | tail -1 > gpurun_out/${t}_config5_n$n.json
cat gpurun_out/${t}_config5_n$n.json
grep -i "via P2P\|NVLS\|Connected all" gpurun_out/${t}_config5_n$n.out | head -12 > gpurun_out/${t}_nccl_p2p.txt
cat gpurun_out/${t}_nccl_p2p.txt | cut -c1-200
rm -f gpurun_out/${t}_config5_n$n.out
tail -c 600 gpurun_out/${t}_config5_n$n.err | tail -4
