#!/usr/bin/env python3
"""Where does the extraction of the config-5 density (768x384x384, sparse surface) spend its time?  Times computeIsosurface_2 with and
without the 16-byte grid_points input (vol_topo), and the CSG extraction of config 2's last step, CUDA events, this library only.

    python tools/topo_probe.py [--steps 5]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import gpucadforam_b200 as g  # noqa: E402
from gpucadforam_b200 import synth  # noqa: E402


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    ctx = g.Context(0, options=0)
    cdims, fdims, d = (384, 192, 192), (768, 384, 384), (0.5, 0.5, 0.5)
    cx, cy, cz = cdims
    fx, fy, fz = fdims
    npts, ncell = fx * fy * fz, (fx - 1) * (fy - 1) * (fz - 1)
    coarse = synth.cantilever_density(cx, cy, cz, struts=40, sigma=1.5, device="cuda").contiguous().reshape(-1)
    lat, iso = g.Gratings(ctx), g.Isosurface(ctx)
    lat.setupTexture(cx, cy, cz)
    pitched_buf = torch.zeros(cx * cy * cz, device="cuda")
    pp = lat.pitched(pitched_buf, cx, cy)
    dens = torch.zeros(npts, device="cuda")
    lat.copytotexture(coarse, pp, cx, cy, cz); lat.updateTexture(pp); lat.refine(dens, fx, fy, fz, *d)
    vol_topo = torch.zeros(npts, 4, dtype=torch.int32, device="cuda")
    result = torch.zeros(npts, device="cuda")
    scr = g.Scratch(ncell)
    probe = g.MeshBuffers(3)
    _, tot0 = iso.computeIsosurface_2(probe.pos, probe.norm, 0.4, scr, fdims, d, (0, 0, 0), 3, vol_topo, dens, 0.0, result)
    mv = tot0 + 3
    mesh = g.MeshBuffers(mv)
    out = {"points": npts, "verts": tot0}
    ms, _ = timed(lambda: iso.computeIsosurface_2(mesh.pos, mesh.norm, 0.4, scr, fdims, d, (0, 0, 0), mv, vol_topo, dens, 0.0, result), args.steps)
    out["topo_with_grid_points_ms"] = ms
    ms, r = timed(lambda: iso.computeIsosurface_2(mesh.pos, mesh.norm, 0.4, scr, fdims, d, (0, 0, 0), mv, None, dens, 0.0, result), args.steps)
    out["topo_without_grid_points_ms"] = ms
    out["verts_without"] = r[1]
    ms, _ = timed(lambda: lat.refine(dens, fx, fy, fz, *d), args.steps)
    out["refine_ms"] = ms
    print(json.dumps(out))


if __name__ == "__main__":
    main()
