#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/tests7.log 2>&1; echo "tests rc=$?" >> $O/tests7.log
tail -8 $O/tests7.log
for i in 1 2; do
timeout 120 python bench.py --profile --fast-field > $O/mc7_new_$i.json 2>&1
GCB_LIB_PATH=$PWD/gpucadforam_b200/libgpucad_b200_base.so timeout 120 python bench.py --profile --fast-field > $O/mc7_base_$i.json 2>&1
done
cat $O/mc7_new_*.json $O/mc7_base_*.json
timeout 300 python tools/config_bench.py > $O/configs7.json 2> $O/configs7.err
SUB="ragged or row_mask or latticeone_three or lattice_variant or csg_pipeline or csg_lattice_modes or topo_three or band_raw or region_three or z_slab or max_verts or empty_and_full"
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "$SUB" 2>&1 | grep -v "^$" | tail -6 > $O/racecheck7.txt
timeout 400 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "$SUB" 2>&1 | grep -v "^$" | tail -6 > $O/synccheck7.txt
cat $O/racecheck7.txt $O/synccheck7.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mc_fused -s 3 -c 1 -f -o $O/r02_mc_fused_v3 python bench.py --profile --fast-field --steps 1 --warmup 3 > $O/ncu_mc7.log 2>&1
