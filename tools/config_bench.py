#!/usr/bin/env python3
"""BASELINE configs 1, 2 and 5 at full size on one B200.  Two arms, kept apart so that bench.py's product arm never loads the
reference library: `ours_*` drive this library through the legacy-signature C ABI (the reference's own call sequences) and,
where one exists, through the fused entry point; `ref_*` drive the reference's own CUDA kernels (oracle/_ref) on the same inputs.
bench.py imports both (`ours` in the default arm, `ref` under --impl reference); run as a script it does both on one GPU, adds
the full-size parity check (counts and whole vertex / normal buffers bit for bit) and prints one JSON line per config.

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

import torch  # noqa: E402

ISO_MASK, BAND_LO, BAND_HI = 0.25, 0.20, 0.30

C1 = dict(n=128, workload="config 1: gyroid TPMS unit lattice 128^3, band [0.20,0.30] (create_lattice, 2 normalises, latticeone)")
C2 = dict(n=256, d=(0.5, 0.5, 0.5), workload="config 2: CSG sphere U box - cylinder on a 256^3 fine grid (3 primitives, 2 retains, computeIsosurface obj_diff)",
          sph=dict(center=(0.0, 0.0, 0.0), radius=40.0, thickness=2.0),
          cub=dict(center=(1.0, 0.5, -0.5), angles=(0.3, 0.2, 0.1), xw=90.0, yw=50.0, zw=60.0),
          cyl=dict(center=(0.0, 0.0, 0.0), axis=(0.0, 0.0, 1.0), radius=18.0, tr=2.0, ta=200.0))
C5 = dict(cdims=(384, 192, 192), fdims=(768, 384, 384), d=(0.5, 0.5, 0.5), iso=0.4,
          workload="config 5: cantilever density 768x384x384 (coarse 384x192x192, 40 struts, sigma 1.5): refine + computeIsosurface_2, iso 0.4")


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


def _res(ms, points, act, tot, **extra):
    out = {"ms": ms, "voxels_per_s": points / (ms * 1e-3), "triangles_per_s": tot / 3 / (ms * 1e-3), "triangles": tot // 3, "active_voxels": act}
    out.update(extra)
    return out


# ------------------------------------------------------------------ config 1
def ours_config1(g, ctx, steps, warmup, keep=None):
    """main.cu:4080-4137: create_lattice -> GPU_buffer_normalise_buffer -> GPU_buffer_normalise_four -> computeIsosurface_latticeone."""
    n = C1["n"]
    dims, npts, ncell = (n, n, n), n ** 3, (n - 1) ** 3
    mv = max(4 * npts, 300000)
    lat, iso = g.Gratings(ctx), g.Isosurface(ctx)
    f, mask, k = (torch.zeros(npts, device="cuda") for _ in range(3))
    scr, mesh = g.Scratch(ncell), g.MeshBuffers(mv)

    def legacy():
        g.Fft_lattice(ctx).create_lattice(f, n, n, n, npts, 0)
        lat.GPU_buffer_normalise_buffer(f, f, npts)
        lat.GPU_buffer_normalise_four(f, mask, k, npts, n, n, n, BAND_LO, BAND_HI)
        return iso.computeIsosurface_latticeone(mask, mesh.pos, mesh.norm, ISO_MASK, scr, dims, (1, 1, 1), (0, 0, 0), mv, k, BAND_LO, BAND_HI)

    ms, (act, tot) = timed(legacy, steps, warmup)
    out = {"workload": C1["workload"], "points": npts, "legacy_calls": _res(ms, npts, act, tot, calls=4)}
    if hasattr(g, "tpms_lattice"):
        mesh_f = g.MeshBuffers(mv)
        ms_f, (a2, t2) = timed(lambda: g.tpms_lattice(ctx, f, 0, dims, ISO_MASK, BAND_LO, BAND_HI, (1, 1, 1), (0, 0, 0), mesh_f.pos, mesh_f.norm, mv)[:2], steps, warmup)
        out["fused_call"] = _res(ms_f, npts, a2, t2, calls=1, same_mesh_as_legacy=bool((a2, t2) == (act, tot) and same(mesh.pos, mesh_f.pos, tot * 16)
                                                                                     and same(mesh.norm, mesh_f.norm, tot * 16)))
    if keep is not None:
        keep.update(k=k, mesh=mesh, counts=(act, tot))
    return out


def ref_config1(g, ref, steps, warmup, keep=None):
    n = C1["n"]
    dims, npts, ncell = (n, n, n), n ** 3, (n - 1) ** 3
    mv = max(4 * npts, 300000)
    f2, mask2, k2 = (torch.zeros(npts, device="cuda") for _ in range(3))
    scr2, mesh2 = g.Scratch(ncell), g.MeshBuffers(mv)

    def theirs():
        ref.create_lattice(f2, n, n, n, 0)
        ref.normalise_buffer(f2, f2, npts)
        ref.normalise_four(f2, mask2, k2, dims, BAND_LO, BAND_HI)
        return ref.isosurface_lattice(True, False, mask2, mesh2.pos, mesh2.norm, ISO_MASK, dims, (1, 1, 1), (0, 0, 0), scr2, mv, k2, None, BAND_LO, BAND_HI)

    ms, (act, tot) = timed(theirs, steps, warmup)
    if keep is not None:
        keep.update(k=k2, mesh=mesh2, counts=(act, tot))
    return {"workload": C1["workload"], "points": npts, "reference_kernels": _res(ms, npts, act, tot)}


# ------------------------------------------------------------------ config 2
def ours_config2(g, ctx, steps, warmup, keep=None):
    """main.cu:3304-3465: sphere -> retain(union); cuboid -> retain(union); cylinder with obj_diff -> mesh."""
    n, d = C2["n"], C2["d"]
    dims, npts, ncell = (n, n, n), n ** 3, (n - 1) ** 3
    mv = max(4 * npts, 300000)
    m, iso = g.Modelling(ctx), g.Isosurface(ctx)
    sph, cub, cyl = C2["sph"], C2["cub"], C2["cyl"]
    zeros = torch.zeros(npts, device="cuda")
    v1, b1, s1, m1 = gp_zeros(npts), torch.zeros(npts, device="cuda"), g.Scratch(ncell), g.MeshBuffers(mv)

    def legacy():
        v1.zero_()
        m.sphere_with_center(b1, sph["center"], sph["radius"], sph["thickness"], n, n, n, *d, False)
        iso.copy_parameter(0.0, dims, d, v1, b1, zeros, obj_union=True)
        m.cuboid(b1, cub["center"], cub["angles"], cub["xw"], cub["yw"], cub["zw"], n, n, n, *d)
        iso.copy_parameter(0.0, dims, d, v1, b1, zeros, obj_union=True)
        m.distance_from_line(b1, cyl["center"], cyl["axis"], cyl["radius"], cyl["tr"], cyl["ta"], n, n, n, *d, False)
        act, tot, _ = iso.computeIsosurface(m1.pos, m1.norm, 0.0, s1, dims, d, (0, 0, 0), mv, v1, b1, zeros, obj_union=False, obj_diff=True)
        return act, tot

    ms, (act, tot) = timed(legacy, steps, warmup)
    out = {"workload": C2["workload"], "points": npts, "legacy_calls": _res(ms, npts, act, tot, calls=6)}
    # the same six calls with GCB_OPT_ASYNC_FIELDS: the five field calls only enqueue, computeIsosurface synchronises once for the counts
    ctx.set_options(g._capi.GCB_OPT_ASYNC_FIELDS)
    try:
        ms_a, (aa, ta) = timed(legacy, steps, warmup)
    finally:
        ctx.set_options(0)
    out["enqueue_only_calls"] = _res(ms_a, npts, aa, ta, calls=6, same_counts_as_blocking=bool((aa, ta) == (act, tot)))
    if hasattr(g, "csg_retain_primitive"):
        v3, b3, m3 = gp_zeros(npts), torch.zeros(npts, device="cuda"), g.MeshBuffers(mv)

        def fused():
            v3.zero_()
            # sphere and cuboid are evaluated inside the retain kernel; their fields are not needed afterwards (b3 receives the cylinder)
            g.csg_retain_primitive(ctx, "sphere", v3, None, dims, d, 0.0, center=sph["center"], radius=sph["radius"], thickness=sph["thickness"])
            g.csg_retain_primitive(ctx, "cuboid", v3, None, dims, d, 0.0, center=cub["center"], angles=cub["angles"], xw=cub["xw"], yw=cub["yw"], zw=cub["zw"])
            m.distance_from_line(b3, cyl["center"], cyl["axis"], cyl["radius"], cyl["tr"], cyl["ta"], n, n, n, *d, False)
            act, tot, _ = iso.computeIsosurface(m3.pos, m3.norm, 0.0, s1, dims, d, (0, 0, 0), mv, v3, b3, zeros, obj_union=False, obj_diff=True)
            return act, tot
        ms_f, (a2, t2) = timed(fused, steps, warmup)
        out["fused_call"] = _res(ms_f, npts, a2, t2, calls=4, same_mesh_as_legacy=bool((a2, t2) == (act, tot) and torch.equal(v1, v3)
                                                                                     and same(m1.pos, m3.pos, tot * 16) and same(m1.norm, m3.norm, tot * 16)))
        ctx.set_options(g._capi.GCB_OPT_ASYNC_FIELDS)
        try:
            ms_fa, (a3, t3) = timed(fused, steps, warmup)
        finally:
            ctx.set_options(0)
        out["fused_enqueue_only_calls"] = _res(ms_fa, npts, a3, t3, calls=4, same_counts_as_blocking=bool((a3, t3) == (act, tot)))
    if keep is not None:
        keep.update(gp=v1, mesh=m1, counts=(act, tot))
    return out


def ref_config2(g, ref, steps, warmup, keep=None):
    n, d = C2["n"], C2["d"]
    dims, npts, ncell = (n, n, n), n ** 3, (n - 1) ** 3
    mv = max(4 * npts, 300000)
    sph, cub, cyl = C2["sph"], C2["cub"], C2["cyl"]
    zeros = torch.zeros(npts, device="cuda")
    v2, b2, s2, m2 = gp_zeros(npts), torch.zeros(npts, device="cuda"), g.Scratch(ncell), g.MeshBuffers(mv)

    def theirs():
        v2.zero_()
        ref.sphere(b2, sph["center"], sph["radius"], sph["thickness"], dims, d, False)
        ref.copy_parameter(v2, b2, zeros, dims, d, 0.0, obj_union=True)
        ref.cuboid(b2, cub["center"], cub["angles"], cub["xw"], cub["yw"], cub["zw"], dims, d)
        ref.copy_parameter(v2, b2, zeros, dims, d, 0.0, obj_union=True)
        ref.distance_from_line(b2, cyl["center"], cyl["axis"], cyl["radius"], cyl["tr"], cyl["ta"], dims, d, False)
        return ref.isosurface_csg(False, m2.pos, m2.norm, 0.0, dims, d, (0, 0, 0), s2, mv, v2, b2, zeros, obj_union=False, obj_diff=True)

    ms, (act, tot) = timed(theirs, steps, warmup)
    if keep is not None:
        keep.update(gp=v2, mesh=m2, counts=(act, tot))
    return {"workload": C2["workload"], "points": npts, "reference_kernels": _res(ms, npts, act, tot)}


# ------------------------------------------------------------------ config 5
def _config5_density():
    from gpucadforam_b200 import synth
    cx, cy, cz = C5["cdims"]
    return synth.cantilever_density(cx, cy, cz, struts=40, sigma=1.5, device="cuda").contiguous().reshape(-1)


def ours_config5(g, ctx, steps, warmup, keep=None):
    """main.cu:3060-3109: refine (2x upsample of the coarse density) + computeIsosurface_2 semantics, iso = VolumeFraction 0.4."""
    cdims, fdims, d = C5["cdims"], C5["fdims"], C5["d"]
    cx, cy, cz = cdims
    fx, fy, fz = fdims
    npts, ncell = fx * fy * fz, (fx - 1) * (fy - 1) * (fz - 1)
    coarse = _config5_density()
    vol_topo = gp_zeros(npts)
    result = torch.zeros(npts, device="cuda")
    lat, iso = g.Gratings(ctx), g.Isosurface(ctx)
    lat.setupTexture(cx, cy, cz)
    pitched_buf = torch.zeros(cx * cy * cz, device="cuda")
    pp = lat.pitched(pitched_buf, cx, cy)
    dens = torch.zeros(npts, device="cuda")
    lat.copytotexture(coarse, pp, cx, cy, cz)
    lat.updateTexture(pp)
    lat.refine(dens, fx, fy, fz, *d)
    scr, probe = g.Scratch(ncell), g.MeshBuffers(3)
    _, tot0 = iso.computeIsosurface_2(probe.pos, probe.norm, C5["iso"], scr, fdims, d, (0, 0, 0), 3, vol_topo, dens, 0.0, result)
    mv = tot0 + 3   # count first, then allocate the mesh exactly (the reference preallocates 4 vertices per point)
    mesh = g.MeshBuffers(mv)

    def legacy():
        lat.copytotexture(coarse, pp, cx, cy, cz)
        lat.updateTexture(pp)
        lat.refine(dens, fx, fy, fz, *d)
        return iso.computeIsosurface_2(mesh.pos, mesh.norm, C5["iso"], scr, fdims, d, (0, 0, 0), mv, vol_topo, dens, 0.0, result)

    ms, (act, tot) = timed(legacy, steps, warmup)
    out = {"workload": C5["workload"], "points": npts, "legacy_calls": _res(ms, npts, act, tot, calls=4)}
    if hasattr(g, "density_surface"):
        mesh_f = g.MeshBuffers(mv)
        ms_f, (a2, t2) = timed(lambda: g.density_surface(ctx, coarse, cdims, dens, fdims, d, C5["iso"], d, (0, 0, 0), mesh_f.pos, mesh_f.norm, mv), steps, warmup)
        out["fused_call"] = _res(ms_f, npts, a2, t2, calls=1, same_mesh_as_legacy=bool((a2, t2) == (act, tot) and same(mesh.pos, mesh_f.pos, tot * 16)
                                                                                     and same(mesh.norm, mesh_f.norm, tot * 16)))
    lat.deleteTexture()
    if keep is not None:
        keep.update(dens=dens, mesh=mesh, counts=(act, tot), mv=mv)
    return out


def ref_config5(g, ref, steps, warmup, keep=None, mv=None):
    cdims, fdims, d = C5["cdims"], C5["fdims"], C5["d"]
    cx, cy, cz = cdims
    fx, fy, fz = fdims
    npts, ncell = fx * fy * fz, (fx - 1) * (fy - 1) * (fz - 1)
    coarse = _config5_density()
    vol_topo = gp_zeros(npts)
    result = torch.zeros(npts, device="cuda")
    dens2 = torch.zeros(npts, device="cuda")
    scr2 = g.Scratch(ncell)
    ref.setup_texture(cx, cy, cz)
    ref.upload_texture(coarse, cx, cy, cz)
    ref.refine(dens2, fdims, d)
    if mv is None:
        probe = g.MeshBuffers(3)
        _, tot0 = ref.isosurface_topo(False, probe.pos, probe.norm, C5["iso"], fdims, d, (0, 0, 0), scr2, 3, vol_topo, dens2, 0.0, result, vol_one=vol_topo, d_solid=dens2)
        mv = tot0 + 3
    mesh2 = g.MeshBuffers(mv)

    def theirs():
        ref.upload_texture(coarse, cx, cy, cz)
        ref.refine(dens2, fdims, d)
        return ref.isosurface_topo(False, mesh2.pos, mesh2.norm, C5["iso"], fdims, d, (0, 0, 0), scr2, mv, vol_topo, dens2, 0.0, result, vol_one=vol_topo, d_solid=dens2)

    ms, (act, tot) = timed(theirs, steps, warmup)
    ref.delete_texture()
    if keep is not None:
        keep.update(dens=dens2, mesh=mesh2, counts=(act, tot))
    return {"workload": C5["workload"], "points": npts, "reference_kernels": _res(ms, npts, act, tot)}


OURS = {"1": ours_config1, "2": ours_config2, "5": ours_config5}
REFS = {"1": ref_config1, "2": ref_config2, "5": ref_config5}


def run_ours(g, ctx, steps, warmup, which=("1", "2", "5")):
    out = {}
    for c in which:
        out[c] = OURS[c](g, ctx, steps, warmup)
        torch.cuda.empty_cache()
    return out


def run_reference(g, ref, steps, warmup, which=("1", "2", "5")):
    out = {}
    for c in which:
        out[c] = REFS[c](g, ref, steps, warmup)
        torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--configs", default="1,2,5")
    ap.add_argument("--tmp", default="/tmp")
    ap.add_argument("--async-fields", action="store_true", help="GCB_OPT_ASYNC_FIELDS: legacy field calls only enqueue (ours arm)")
    args = ap.parse_args()
    import gpucadforam_b200 as g
    import ref_py as ref
    if not ref.available():
        print(json.dumps({"error": "oracle/_ref/libgpucad_ref.so not built"}))
        return 1
    ctx = g.Context(0, options=g._capi.GCB_OPT_ASYNC_FIELDS if args.async_fields else 0)
    for c in [x.strip() for x in args.configs.split(",")]:
        ko, kr = {}, {}
        o = OURS[c](g, ctx, args.steps, args.warmup, keep=ko)
        r = REFS[c](g, ref, args.steps, args.warmup, keep=kr, **({"mv": ko["mv"]} if c == "5" else {}))
        tot = ko["counts"][1]
        parity = {"counts": ko["counts"] == kr["counts"], "pos_bits": same(ko["mesh"].pos, kr["mesh"].pos, tot * 16),
                  "norm_bits": same(ko["mesh"].norm, kr["mesh"].norm, tot * 16)}
        for key in ("k", "gp", "dens"):
            if key in ko:
                parity[key + "_bits"] = bool(torch.equal(ko[key].view(torch.uint8), kr[key].view(torch.uint8)))
        line = {"config": int(c)}
        line.update(o)
        line.update(reference_kernels=r["reference_kernels"], parity_full_size=parity,
                    speedup_legacy_calls=r["reference_kernels"]["ms"] / o["legacy_calls"]["ms"],
                    legacy_calls_mode="enqueue only (GCB_OPT_ASYNC_FIELDS)" if args.async_fields else "blocking (as the reference wrappers)")
        if "fused_call" in o:
            line["speedup_fused_call"] = r["reference_kernels"]["ms"] / o["fused_call"]["ms"]
        if "enqueue_only_calls" in o:
            line["speedup_enqueue_only_calls"] = r["reference_kernels"]["ms"] / o["enqueue_only_calls"]["ms"]
        if "fused_enqueue_only_calls" in o:
            line["speedup_fused_enqueue_only_calls"] = r["reference_kernels"]["ms"] / o["fused_enqueue_only_calls"]["ms"]
        if c == "2":
            p1, p2 = os.path.join(args.tmp, "ours.obj"), os.path.join(args.tmp, "ref.obj")
            t0 = time.time(); g.File_output(ctx).file_write_obj(ko["mesh"].pos, tot, p1); t_o = time.time() - t0
            t0 = time.time(); ref.write_obj(kr["mesh"].pos, tot, p2); t_r = time.time() - t0
            line["obj_export"] = {"ours_s": t_o, "reference_writer_s": t_r, "bytes": os.path.getsize(p1), "same_bytes": open(p1, "rb").read() == open(p2, "rb").read()}
        print(json.dumps(line), flush=True)
        del ko, kr
        torch.cuda.empty_cache()
    return 0


if __name__ == "__main__":
    sys.exit(main())
