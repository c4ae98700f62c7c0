#!/usr/bin/env python
"""Profiling driver for the SPARSE-surface case of the extraction kernel: BASELINE config 5's fused call (gcb_density_surface, 768 x 384 x 384)
run a few times and nothing else, so that `ncu -k regex:mc_fused -s 2 -c 1` lands on a warm launch.  Prints the call time."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import gpucadforam_b200 as g
import config_bench as cb

ctx = g.Context(0)
cdims, fdims, d = cb.C5["cdims"], cb.C5["fdims"], cb.C5["d"]
npts = fdims[0] * fdims[1] * fdims[2]
coarse = cb._config5_density()
dens = torch.zeros(npts, device="cuda")
probe = g.MeshBuffers(3)
_, tot = g.density_surface(ctx, coarse, cdims, dens, fdims, d, cb.C5["iso"], d, (0, 0, 0), probe.pos, probe.norm, 3)
mv = tot + 3
mesh = g.MeshBuffers(mv)
for i in range(int(os.environ.get("REPS", "4"))):
    torch.cuda.synchronize()
    t = time.perf_counter()
    a, v = g.density_surface(ctx, coarse, cdims, dens, fdims, d, cb.C5["iso"], d, (0, 0, 0), mesh.pos, mesh.norm, mv)
    torch.cuda.synchronize()
    print("call %d: %.3f ms  active %d verts %d" % (i, (time.perf_counter() - t) * 1e3, a, v))
