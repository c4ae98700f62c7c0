#!/bin/bash
# Round evidence on one B200: GPU parity tests, both bench arms, the legacy-sequence configs, the ncu launch list of the bench
# command and one `ncu --set full` capture of each of the two kernels of the step.  Writes into gpurun_out/ (tag = $1).
#   gpurun --timeout 1500 -- 'bash tools/evidence.sh r01b'
t=${1:-ev}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/${t}_tests.log
tail -2 gpurun_out/${t}_tests.log
timeout 300 python bench.py --impl reference 2>gpurun_out/${t}_bench_reference.err | tail -1 > gpurun_out/${t}_bench_reference.json
timeout 300 python bench.py 2>gpurun_out/${t}_bench_ours.err | tail -1 > gpurun_out/${t}_bench_ours.json
cat gpurun_out/${t}_bench_ours.json
(timeout 300 python tools/config_bench.py 2>&1 | tail -3) > gpurun_out/${t}_configs.json
timeout 200 python tools/topo_probe.py > gpurun_out/${t}_topo_probe.json 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"svl_|mc_fused|minmax" -c 60 --csv --log-file gpurun_out/${t}_launches.csv python bench.py --steps 2 --warmup 3 --profile > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mc_fused -s 4 -c 1 -f -o gpurun_out/${t}_mc python bench.py --steps 2 --warmup 3 --profile > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:svl_field -s 4 -c 1 -f -o gpurun_out/${t}_field python bench.py --steps 2 --warmup 3 --profile > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mc_fused -s 4 -c 1 -f -o gpurun_out/${t}_topo python tools/topo_probe.py --steps 1 > /dev/null 2>&1
ls -la gpurun_out | grep ${t}_
# C++ host side and the enqueue-only legacy sequences
for c in 1 2 3 5; do timeout 60 ./gpucadforam_b200/gpucad_headless $c 2>&1 | tail -1; done > gpurun_out/${t}_headless.txt
(timeout 300 python tools/config_bench.py --async-fields 2>&1 | tail -3) > gpurun_out/${t}_configs_async.json
timeout 200 python tools/config5_multi.py --check 2>/dev/null | tail -1 > gpurun_out/${t}_config5_n1.json
# multi-GPU (separate calls, charged N x):  gpurun --gpus 8 -- 'bash tools/_scale.sh'   and   gpurun --gpus 8 -- 'bash tools/halo_run.sh TAG 8 only'
