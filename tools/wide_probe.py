#!/usr/bin/env python3
"""Extraction (fused band-raw mode) on a slab with 2048-wide rows -- the per-rank shape of BASELINE config 4 (2048^3 on 8 GPUs is
2048 x 2048 x 257 per rank; this probe takes a thinner slab so it is quick).  Tile height is limited by shared memory there, so the
tile-size knobs matter: GCB_MC_SMEM_CAP_KB / GCB_MC_TILE_CELLS.    python tools/wide_probe.py [--nx 2048 --nz 65]"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import gpucadforam_b200 as g  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=2048)
    ap.add_argument("--nz", type=int, default=65)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    nx = ny = a.nx
    nz = a.nz
    ctx = g.Context(0, options=0)
    x = torch.arange(nx, device="cuda", dtype=torch.float32) * (2 * math.pi / 40.0)
    z = torch.arange(nz, device="cuda", dtype=torch.float32) * (2 * math.pi / 40.0)
    f = (torch.cos(x)[None, None, :] * torch.sin(x)[None, :, None] + torch.cos(x)[None, :, None] * torch.sin(z)[:, None, None]
         + torch.cos(z)[:, None, None] * torch.sin(x)[None, None, :]).contiguous().reshape(-1)
    lo, hi = float(f.min()), float(f.max())
    dims = (nx, ny, nz)
    act, tot = g.extract_band_raw(ctx, f, lo, hi, 0.25, 0.2, 0.3, dims, (0.25, 0.25, 0.25), (0, 0, 0), None, None, 0, count_only=True)
    mesh = g.MeshBuffers(tot + 3)

    def run():
        return g.extract_band_raw(ctx, f, lo, hi, 0.25, 0.2, 0.3, dims, (0.25, 0.25, 0.25), (0, 0, 0), mesh.pos, mesh.norm, tot + 3)

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        r = run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    pts = nx * ny * nz
    print(json.dumps({"dims": dims, "verts": r[1], "ms": ms, "gvoxel_per_s": pts / ms / 1e6, "hbm_frac": (4.0 * pts + 32.0 * r[1]) / (ms * 1e-3) / 6454.9e9,
                      "cap_kb": os.environ.get("GCB_MC_SMEM_CAP_KB"), "tile_cells": os.environ.get("GCB_MC_TILE_CELLS")}))


if __name__ == "__main__":
    main()
