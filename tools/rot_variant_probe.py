"""Experiment: which explicit-fma spelling of the Euler rotation reproduces the reference primitives bit for bit?
Runs every variant (GCB_ROT_VARIANT bit field, read per launch) of the rotated primitives against the reference kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gpucadforam_b200 as g
import ref_py as ref
ctx = g.Context(0, options=0)
m = g.Modelling(ctx)
dims, d = (40, 36, 44), (0.5, 0.5, 0.5)
nx, ny, nz = dims
n = nx * ny * nz
cases = []
for center, ang in (((0.7, -0.4, 0.3), (0.3, 0.2, 0.1)), ((-1.1, 0.2, 0.9), (1.3, -0.7, 2.4)), ((0, 0, 0), (0.0, 0.5, 0.0))):
    cases += [("cuboid", lambda o, c=center, a=ang: m.cuboid(o, c, a, 11.0, 9.0, 7.0, nx, ny, nz, *d), lambda o, c=center, a=ang: ref.cuboid(o, c, a, 11.0, 9.0, 7.0, dims, d)),
              ("torus", lambda o, c=center, a=ang: m.torus_with_center(o, c, a, 6.0, 2.0, nx, ny, nz, *d), lambda o, c=center, a=ang: ref.torus(o, c, a, 6.0, 2.0, dims, d)),
              ("cone", lambda o, c=center, a=ang: m.cone_with_base_radius_height(o, c, a, 5.0, 8.0, nx, ny, nz, *d), lambda o, c=center, a=ang: ref.cone(o, c, a, 5.0, 8.0, dims, d)),
              ("cone_frustum", lambda o, c=center, a=ang: m.cone_frustum(o, c, a, 2.0, 5.0, 8.0, nx, ny, nz, *d), lambda o, c=center, a=ang: ref.cone_frustum(o, c, a, 2.0, 5.0, 8.0, dims, d)),
              ("pyramid", lambda o, c=center, a=ang: m.pyramid_frustum(o, c, a, 9.0, 4.0, 8.0, 7.0, 3.0, nx, ny, nz, *d), lambda o, c=center, a=ang: ref.pyramid_frustum(o, c, a, 9.0, 4.0, 8.0, 7.0, 3.0, dims, d)),
              ("shell", lambda o, c=center, a=ang: m.cuboid_shell(o, c, a, 11.0, 9.0, 7.0, 1.0, nx, ny, nz, *d), lambda o, c=center, a=ang: ref.cuboid_shell(o, c, a, 11.0, 9.0, 7.0, 1.0, dims, d))]
refs = []
for name, mine, theirs in cases:
    o = torch.zeros(n, device="cuda"); theirs(o); refs.append(o.view(torch.int32).clone())
for v in [-1] + list(range(64)):
    os.environ["GCB_ROT_VARIANT"] = str(v)
    bad = {}
    for (name, mine, theirs), r in zip(cases, refs):
        o = torch.zeros(n, device="cuda"); mine(o)
        bad[name] = bad.get(name, 0) + int((o.view(torch.int32) != r).sum())
    print(v, bad, flush=True)
