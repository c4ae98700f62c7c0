"""Second stage: many random angle sets on the cuboid / torus for the candidate variants."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import gpucadforam_b200 as g
import ref_py as ref
ctx = g.Context(0, options=0)
m = g.Modelling(ctx)
dims, d = (24, 20, 28), (0.5, 0.5, 0.5)
nx, ny, nz = dims
n = nx * ny * nz
rng = np.random.RandomState(5)
sets = [(tuple(float(np.float32(v)) for v in rng.uniform(-2, 2, 3)), tuple(float(np.float32(v)) for v in rng.uniform(-3.2, 3.2, 3))) for _ in range(60)]
refs = []
for c, a in sets:
    o = torch.zeros(n, device="cuda"); ref.cuboid(o, c, a, 7.0, 5.0, 6.0, dims, d); refs.append(o.view(torch.int32).clone())
for v in (-1, 9, 11, 13, 15, 25, 41):
    os.environ["GCB_ROT_VARIANT"] = str(v)
    bad = 0; badsets = 0
    for (c, a), r in zip(sets, refs):
        o = torch.zeros(n, device="cuda"); m.cuboid(o, c, a, 7.0, 5.0, 6.0, nx, ny, nz, *d)
        b = int((o.view(torch.int32) != r).sum()); bad += b; badsets += b > 0
    print(v, bad, badsets, flush=True)
