#!/usr/bin/env python3
"""SVL phase solve at the bench's control-grid size: the reference's per-harmonic loop
(finding_phi + GPUCG_lattice, host-driven CG, main.cu:3949-3962) next to gcb_svl_phase_solve (all harmonics batched, CG scalars on
the device).  Wall clock around each, results compared bit for bit.

    python tools/phase_solve_bench.py [--control 128] > profiles/rNN_phase_solve.json
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import gpucadforam_b200 as g  # noqa: E402
from gpucadforam_b200 import synth  # noqa: E402
import ref_py as ref  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--control", type=int, default=128)
    ap.add_argument("--iters", type=int, default=500)
    ap.add_argument("--end-res", type=float, default=0.01)
    args = ap.parse_args()
    c = args.control
    dims, d = (c, c, c), (1.0, 1.0, 1.0)
    n = c ** 3
    harm = synth.HARMONICS
    # period field as the reference builds it: period_data + GPU_buffer_normalise_three(NumX/10, NumX/4) (main.cu:3927-3931)
    period = torch.zeros(n, device="cuda")
    have_ref = ref.available()
    if have_ref:
        ref.period_data(period, dims, d, (c / 2.0,) * 3, "z")
        ref.normalise_three(period, period, n, float(c // 10), float(c // 4))
    else:
        zz, yy, xx = np.meshgrid(np.arange(c), np.arange(c), np.arange(c), indexing="ij")
        r = np.sqrt((xx - c / 2.0 + 1) ** 2 + (yy - c / 2.0 + 1) ** 2).astype(np.float32)
        period = torch.tensor((c // 10 + (c // 4) * (r - r.min()) / (r.max() - r.min())).astype(np.float32).reshape(-1), device="cuda")
    ctx = g.Context(0, options=0)
    phi = torch.zeros(len(harm), n, device="cuda")
    g.svl_phase_solve(ctx, phi, period, harm[:2], dims, d, iters=5)   # warm-up (module load, allocations)
    torch.cuda.synchronize()
    t0 = time.time()
    fi, fr = g.svl_phase_solve(ctx, phi, period, harm, dims, d, latticetype="r", uniform_type=2, iters=args.iters, end_res=args.end_res)
    torch.cuda.synchronize()
    t_ours = time.time() - t0
    out = {"control_grid": list(dims), "harmonics": len(harm), "iters_cap": args.iters, "end_res": args.end_res, "ours_s": t_ours,
           "cg_iterations_total": int(sum(fi) - len(fi)), "cg_iterations_max": int(max(fi)) - 1}
    if have_ref:
        rphi = torch.zeros(n, device="cuda")
        ref.finding_phi(rphi, period, dims, harm[0], d, latticetype="r", uniform_type=2)
        ref.cg(rphi, dims, 5, args.end_res)   # warm-up
        torch.cuda.synchronize()
        t0 = time.time()
        same, iters_equal = True, True
        for hi, h in enumerate(harm):
            ref.finding_phi(rphi, period, dims, h, d, latticetype="r", uniform_type=2)
            fi_r, fr_r = ref.cg(rphi, dims, args.iters, args.end_res)
            if hi < 8 or hi % 9 == 0:   # bit comparison on a sample of harmonics inside the loop (costs a few ms each)
                same = same and bool(torch.equal(rphi.view(torch.int32), phi[hi].view(torch.int32)))
            iters_equal = iters_equal and fi_r == fi[hi] and fr_r == fr[hi]
        torch.cuda.synchronize()
        t_ref = time.time() - t0
        out.update({"reference_loop_s": t_ref, "speedup": t_ref / t_ours, "bit_identical_solutions_sampled": same, "iterations_and_residuals_equal": iters_equal})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
