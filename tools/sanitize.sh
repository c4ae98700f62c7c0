#!/bin/bash
# compute-sanitizer over the GPU parity tests (one B200).  memcheck: everything but the large grids; racecheck / synccheck: the tests
# that drive the fused extraction kernel through all of its modes and paths.   gpurun --timeout 1200 -- 'bash tools/sanitize.sh'
mkdir -p gpurun_out
out=gpurun_out/compute_sanitizer.txt
SUB="ragged or row_mask or latticeone_three or lattice_variant or csg_pipeline or csg_lattice_modes or topo_three or band_raw or region_three or z_slab or max_verts or empty_and_full"
{
echo "# compute-sanitizer runs over the GPU parity tests on one B200 (tools/sanitize.sh)"
echo; echo "## memcheck (parity suite): compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k 'not large_grid and not bench_like and not obj_device'"
timeout 700 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "not large_grid and not bench_like and not obj_device" 2>&1 | grep -v "^$" | tail -4
echo; echo "## memcheck (full-size BASELINE configs, own process: the two suites together exhaust the tool's tracking under one process): compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_fullsize.py -m gpu -q"
timeout 500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_fullsize.py -m gpu -q 2>&1 | grep -v "^$" | tail -4
echo; echo "## racecheck: compute-sanitizer --tool racecheck python -m pytest tests -m gpu -q -k '$SUB'"
timeout 500 compute-sanitizer --tool racecheck python -m pytest tests -m gpu -q -k "$SUB" 2>&1 | grep -v "^$" | tail -6
echo; echo "## synccheck: compute-sanitizer --tool synccheck python -m pytest tests -m gpu -q -k '$SUB'"
timeout 300 compute-sanitizer --tool synccheck python -m pytest tests -m gpu -q -k "$SUB" 2>&1 | grep -v "^$" | tail -6
} > $out 2>&1
cat $out
