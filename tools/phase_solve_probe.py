"""Experiment / check: finding_phi contraction variants and the batched CG against the reference kernels (bit for bit)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import gpucadforam_b200 as g
import ref_py as ref
ctx = g.Context(0, options=0)
dims, d = (20, 16, 12), (1.0, 1.0, 1.0)
nx, ny, nz = dims
n = nx * ny * nz
rng = np.random.RandomState(3)
period = torch.tensor(rng.uniform(2.0, 5.0, n).astype(np.float32), device="cuda")
cases = [dict(latticetype=t, uniform_type=u) for t in "rbn" for u in (2, 0, 1)] + [dict(latticetype="s", uniform_type=2), dict(latticetype="s", uniform_type=2, sinewave_zaxis=True)]
harm = [(1, 0, 0), (0, 1, 0), (0, 0, 1), (1, -2, 1), (-2, 1, 2), (2, 2, -1)]
kw = dict(const_period=7.3, periods=(6.1, 7.7, 5.3), lcon=0.45, lcon_1=0.07)
refs = {}
for ci, cs in enumerate(cases):
    for h in harm:
        o = torch.zeros(n, device="cuda"); ref.finding_phi(o, period, dims, h, d, **cs, **kw); refs[(ci, h)] = o.view(torch.int32).clone()
for v in range(32):
    os.environ["GCB_PHI_VARIANT"] = str(v)
    bad = [0] * len(cases)
    for ci, cs in enumerate(cases):
        for h in harm:
            o = torch.zeros(n, device="cuda"); g.finding_phi(ctx, o, period, dims, h, d, **cs, **kw)
            bad[ci] += int((o.view(torch.int32) != refs[(ci, h)]).sum())
    print("variant", v, bad, flush=True)
    if sum(bad) == 0: break
# CG: reference per harmonic vs ours (single and batched)
dims2 = (32, 32, 16)
n2 = dims2[0] * dims2[1] * dims2[2]
period2 = torch.tensor(rng.uniform(3.0, 8.0, n2).astype(np.float32), device="cuda")
allphi = torch.zeros(len(harm), n2, device="cuda")
fi_b, fr_b = g.svl_phase_solve(ctx, allphi, period2, harm, dims2, d, latticetype="r", uniform_type=2, iters=500, end_res=0.01)
for hi, h in enumerate(harm):
    r = torch.zeros(n2, device="cuda"); ref.finding_phi(r, period2, dims2, h, d, latticetype="r", uniform_type=2)
    m = r.clone()
    fi_r, fr_r = ref.cg(r, dims2, 500, 0.01)
    fi_m, fr_m = g.GPUCG_lattice(ctx, m, dims2, 500, 0.01)
    print(h, "ref iters", fi_r, fr_r, "| single", fi_m, fr_m, "bits differ", int((m.view(torch.int32) != r.view(torch.int32)).sum()),
          "| batched", fi_b[hi], fr_b[hi], "bits differ", int((allphi[hi].view(torch.int32) != r.view(torch.int32)).sum()), flush=True)
