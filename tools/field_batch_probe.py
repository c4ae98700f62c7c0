"""Timing probe: field kernel launch cost versus number of harmonics per launch (device-resident control grids)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpucadforam_b200 as g
from gpucadforam_b200 import synth
F, R = 512, 4
dev = torch.device("cuda", 0)
coef = synth.gyroid_coefficients()
harm = synth.HARMONICS
c = F // R
phi = synth.phase_grids(c, c, c, device=dev, z0=0, cz_total=c, harmonics=harm, periods=F / 40.0)
ctx = g.Context(0, options=0)
svl = torch.zeros(F * F * F, device=dev)
d = (1.0 / R,) * 3
for nh in (1, 2, 4, 8, 13, 16, 31, 62):
    for acc in (False, True):
        ts = []
        for it in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            g.svl_field(ctx, svl, phi[:nh].contiguous(), coef[:nh], (c, c, c), (F, F, F), d, accumulate=acc)
            e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        print(f"nh={nh:3d} accumulate={int(acc)}  {min(ts):.3f} ms   per harmonic {min(ts)/nh:.4f}")
