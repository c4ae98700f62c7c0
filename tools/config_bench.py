#!/usr/bin/env python3
"""BASELINE configs 1, 2 and 5 at full size on one B200: this library (legacy-signature C ABI) next to the reference's own
CUDA kernels (oracle/_ref), same inputs, CUDA-event timing, and a full-size parity check of the results (counts and whole
vertex / normal buffers bit for bit).  Config 3 is bench.py's default line, config 4 its --gpus 8 --strong line.

    python tools/config_bench.py [--steps 5] [--warmup 3] > profiles/rNN_configs.json
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

ISO_MASK, BAND_LO, BAND_HI = 0.25, 0.20, 0.30
PEAK_GBS = 6456.2
LEGACY_CALLS = "blocking (as the reference wrappers)"


def timed(fn, steps, warmup):
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


def gp_zeros(n):
    return torch.zeros(n, 4, dtype=torch.int32, device="cuda")


def same(a, b, nbytes):
    return bool(torch.equal(a.view(torch.uint8).reshape(-1)[:nbytes], b.view(torch.uint8).reshape(-1)[:nbytes]))


def line(cfg, desc, points, ms_ours, ms_ref, act, tot, parity, alg_bytes, extra=None):
    out = {"config": cfg, "workload": desc, "points": points, "active_voxels": act, "triangles": tot // 3,
           "ours": {"ms": ms_ours, "voxels_per_s": points / (ms_ours * 1e-3), "triangles_per_s": tot / 3 / (ms_ours * 1e-3)},
           "reference_kernels": {"ms": ms_ref, "voxels_per_s": points / (ms_ref * 1e-3)},
           "speedup_vs_reference_kernels": ms_ref / ms_ours, "parity_full_size": parity,
           "algorithmic_bytes": alg_bytes, "hbm_roofline_frac_whole_step": alg_bytes / (ms_ours * 1e-3) / 1e9 / PEAK_GBS}
    out["legacy_calls"] = LEGACY_CALLS
    if extra:
        out.update(extra)
    print(json.dumps(out), flush=True)


def config1(ctx, args):
    """Gyroid TPMS unit lattice 128^3: create_lattice -> normalise_buffer -> normalise_four -> latticeone (main.cu:4080-4137)."""
    n = 128
    dims, npts, ncell = (n, n, n), n ** 3, (n - 1) ** 3
    mv = max(4 * npts, 300000)
    lat, iso = g.Gratings(ctx), g.Isosurface(ctx)
    f, mask, k = (torch.zeros(npts, device="cuda") for _ in range(3))
    scr, mesh = g.Scratch(ncell), g.MeshBuffers(mv)

    def ours():
        g.Fft_lattice(ctx).create_lattice(f, n, n, n, npts, 0)
        lat.GPU_buffer_normalise_buffer(f, f, npts)
        lat.GPU_buffer_normalise_four(f, mask, k, npts, n, n, n, BAND_LO, BAND_HI)
        return iso.computeIsosurface_latticeone(mask, mesh.pos, mesh.norm, ISO_MASK, scr, dims, (1, 1, 1), (0, 0, 0), mv, k, BAND_LO, BAND_HI)

    f2, mask2, k2 = (torch.zeros(npts, device="cuda") for _ in range(3))
    scr2, mesh2 = g.Scratch(ncell), g.MeshBuffers(mv)

    def theirs():
        ref.create_lattice(f2, n, n, n, 0)
        ref.normalise_buffer(f2, f2, npts)
        ref.normalise_four(f2, mask2, k2, dims, BAND_LO, BAND_HI)
        return ref.isosurface_lattice(True, False, mask2, mesh2.pos, mesh2.norm, ISO_MASK, dims, (1, 1, 1), (0, 0, 0), scr2, mv, k2, None, BAND_LO, BAND_HI)

    ms_o, (act, tot) = timed(ours, args.steps, args.warmup)
    ms_r, (a2, t2) = timed(theirs, args.steps, args.warmup)
    parity = {"counts": (act, tot) == (a2, t2), "field_bits": same(k, k2, npts * 4), "pos_bits": same(mesh.pos, mesh2.pos, tot * 16),
              "norm_bits": same(mesh.norm, mesh2.norm, tot * 16)}
    # stored-field formulation the reference imposes: field written once, read by two normalise passes, mask + k written and read
    line(1, "gyroid TPMS unit lattice 128^3, band [0.20,0.30], legacy call sequence (5 calls)", npts, ms_o, ms_r, act, tot, parity,
         4.0 * npts * 6 + 32.0 * tot)


def config2(ctx, args, tmpdir):
    """CSG sphere U box - cylinder on a 256^3 fine grid (dx2 = 0.5) + .obj export (main.cu:3304-3465, :4695-4778)."""
    n = 256
    dims, d, npts, ncell = (n, n, n), (0.5, 0.5, 0.5), n ** 3, (n - 1) ** 3
    mv = max(4 * npts, 300000)
    m, iso = g.Modelling(ctx), g.Isosurface(ctx)
    sph = dict(center=(0.0, 0.0, 0.0), radius=40.0, thickness=2.0)
    cub = dict(center=(1.0, 0.5, -0.5), angles=(0.3, 0.2, 0.1), xw=90.0, yw=50.0, zw=60.0)
    cyl = dict(center=(0.0, 0.0, 0.0), axis=(0.0, 0.0, 1.0), radius=18.0, tr=2.0, ta=200.0)
    zeros = torch.zeros(npts, device="cuda")

    def run(use_ref, vol_one, boundary, scr, mesh):
        vol_one.zero_()
        if use_ref:
            ref.sphere(boundary, sph["center"], sph["radius"], sph["thickness"], dims, d, False)
            ref.copy_parameter(vol_one, boundary, zeros, dims, d, 0.0, obj_union=True)
            ref.cuboid(boundary, cub["center"], cub["angles"], cub["xw"], cub["yw"], cub["zw"], dims, d)
            ref.copy_parameter(vol_one, boundary, zeros, dims, d, 0.0, obj_union=True)
            ref.distance_from_line(boundary, cyl["center"], cyl["axis"], cyl["radius"], cyl["tr"], cyl["ta"], dims, d, False)
            return ref.isosurface_csg(False, mesh.pos, mesh.norm, 0.0, dims, d, (0, 0, 0), scr, mv, vol_one, boundary, zeros, obj_union=False, obj_diff=True)
        m.sphere_with_center(boundary, sph["center"], sph["radius"], sph["thickness"], n, n, n, *d, False)
        iso.copy_parameter(0.0, dims, d, vol_one, boundary, zeros, obj_union=True)
        m.cuboid(boundary, cub["center"], cub["angles"], cub["xw"], cub["yw"], cub["zw"], n, n, n, *d)
        iso.copy_parameter(0.0, dims, d, vol_one, boundary, zeros, obj_union=True)
        m.distance_from_line(boundary, cyl["center"], cyl["axis"], cyl["radius"], cyl["tr"], cyl["ta"], n, n, n, *d, False)
        act, tot, _ = iso.computeIsosurface(mesh.pos, mesh.norm, 0.0, scr, dims, d, (0, 0, 0), mv, vol_one, boundary, zeros, obj_union=False, obj_diff=True)
        return act, tot

    v1, b1, s1, m1 = gp_zeros(npts), torch.zeros(npts, device="cuda"), g.Scratch(ncell), g.MeshBuffers(mv)
    v2, b2, s2, m2 = gp_zeros(npts), torch.zeros(npts, device="cuda"), g.Scratch(ncell), g.MeshBuffers(mv)
    ms_o, (act, tot) = timed(lambda: run(False, v1, b1, s1, m1), args.steps, args.warmup)
    ms_r, (a2, t2) = timed(lambda: run(True, v2, b2, s2, m2), args.steps, args.warmup)
    parity = {"counts": (act, tot) == (a2, t2), "grid_points_bits": same(v1, v2, npts * 16), "pos_bits": same(m1.pos, m2.pos, tot * 16),
              "norm_bits": same(m1.norm, m2.norm, tot * 16)}
    # .obj export: ours (hash weld) vs the reference writer (std::map weld), wall clock, same bytes
    p1, p2 = os.path.join(tmpdir, "ours.obj"), os.path.join(tmpdir, "ref.obj")
    t0 = time.time(); g.File_output(ctx).file_write_obj(m1.pos, tot, p1); t_obj_o = time.time() - t0
    t0 = time.time(); ref.write_obj(m2.pos, t2, p2); t_obj_r = time.time() - t0
    parity["obj_bytes"] = open(p1, "rb").read() == open(p2, "rb").read()
    # 3 primitive fields written (4 B), 2 retains (grid_points 16 R + 16 W, field 4 R + 3 neighbours cached), extraction 24 B/pt
    line(2, "CSG sphere U box - cylinder, 256^3 fine grid: 3 primitives, 2 retains, computeIsosurface (obj_diff)", npts, ms_o, ms_r, act, tot, parity,
         npts * (3 * 4.0 + 2 * 36.0 + 24.0) + 32.0 * tot,
         {"obj_export": {"ours_s": t_obj_o, "reference_writer_s": t_obj_r, "bytes": os.path.getsize(p1), "note": "host-side weld + text, wall clock, 1 thread each"}})


def config5(ctx, args):
    """Synthetic cantilever density 768x384x384: refine (2x upsample) + computeIsosurface_2 semantics, iso = 0.4 (main.cu:3060-3109)."""
    cdims, fdims, d = (384, 192, 192), (768, 384, 384), (0.5, 0.5, 0.5)
    cx, cy, cz = cdims
    fx, fy, fz = fdims
    npts, ncell = fx * fy * fz, (fx - 1) * (fy - 1) * (fz - 1)
    coarse = synth.cantilever_density(cx, cy, cz, struts=40, sigma=1.5, device="cuda").contiguous().reshape(-1)
    vol_topo = gp_zeros(npts)
    result = torch.zeros(npts, device="cuda")
    lat, iso = g.Gratings(ctx), g.Isosurface(ctx)
    lat.setupTexture(cx, cy, cz)
    pitched_buf = torch.zeros(cx * cy * cz, device="cuda")
    pp = lat.pitched(pitched_buf, cx, cy)
    dens, dens2 = torch.zeros(npts, device="cuda"), torch.zeros(npts, device="cuda")
    # count first, then allocate the mesh exactly (the reference preallocates 4 vertices per point)
    lat.copytotexture(coarse, pp, cx, cy, cz); lat.updateTexture(pp); lat.refine(dens, fx, fy, fz, *d)
    scr = g.Scratch(ncell)
    probe = g.MeshBuffers(3)
    act0, tot0 = iso.computeIsosurface_2(probe.pos, probe.norm, 0.4, scr, fdims, d, (0, 0, 0), 3, vol_topo, dens, 0.0, result)
    mv = tot0 + 3
    mesh, mesh2, scr2 = g.MeshBuffers(mv), g.MeshBuffers(mv), g.Scratch(ncell)

    def ours():
        lat.copytotexture(coarse, pp, cx, cy, cz)
        lat.updateTexture(pp)
        lat.refine(dens, fx, fy, fz, *d)
        return iso.computeIsosurface_2(mesh.pos, mesh.norm, 0.4, scr, fdims, d, (0, 0, 0), mv, vol_topo, dens, 0.0, result)

    ref.setup_texture(cx, cy, cz)

    def theirs():
        ref.upload_texture(coarse, cx, cy, cz)
        ref.refine(dens2, fdims, d)
        return ref.isosurface_topo(False, mesh2.pos, mesh2.norm, 0.4, fdims, d, (0, 0, 0), scr2, mv, vol_topo, dens2, 0.0, result, vol_one=vol_topo, d_solid=dens2)

    ms_o, (act, tot) = timed(ours, args.steps, args.warmup)
    ms_r, (a2, t2) = timed(theirs, args.steps, args.warmup)
    ref.delete_texture()
    parity = {"counts": (act, tot) == (a2, t2), "density_bits": same(dens, dens2, npts * 4), "pos_bits": same(mesh.pos, mesh2.pos, tot * 16),
              "norm_bits": same(mesh.norm, mesh2.norm, tot * 16)}
    # refine writes 4 B/pt; extraction reads density 4 + grid_points 16 + d_result 4 (only at active cells) per point
    line(5, "cantilever density 768x384x384 (coarse 384x192x192, 40 struts, sigma 1.5): refine + computeIsosurface_2, iso 0.4", npts, ms_o, ms_r, act, tot,
         parity, npts * (4.0 + 20.0) + 32.0 * tot)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--configs", default="1,2,5")
    ap.add_argument("--tmp", default="/tmp")
    ap.add_argument("--async-fields", action="store_true", help="GCB_OPT_ASYNC_FIELDS: legacy field calls only enqueue (ours arm)")
    args = ap.parse_args()
    if not ref.available():
        print(json.dumps({"error": "oracle/_ref/libgpucad_ref.so not built"}))
        return 1
    global LEGACY_CALLS
    if args.async_fields:
        LEGACY_CALLS = "ours: enqueue only (GCB_OPT_ASYNC_FIELDS); reference kernels: blocking"
    ctx = g.Context(0, options=g._capi.GCB_OPT_ASYNC_FIELDS if args.async_fields else 0)
    for c in args.configs.split(","):
        {"1": lambda: config1(ctx, args), "2": lambda: config2(ctx, args, args.tmp), "5": lambda: config5(ctx, args)}[c.strip()]()
    return 0


if __name__ == "__main__":
    sys.exit(main())
