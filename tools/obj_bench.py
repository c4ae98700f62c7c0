#!/usr/bin/env python3
"""`.obj` export throughput: device writer (default) vs host restatement (GCB_OPT_OBJ_HOST) vs the reference's
own writer (oracle/_ref, std::map weld) on prefixes of the bench lattice mesh (config 3, 512^3 -> 172.6 M vertices).
Wall-clock seconds including the file write to --dir; the byte-equality of the three files is asserted where more than one ran.

    python tools/obj_bench.py [--fine 512] [--dir /tmp] > profiles/rNN_obj_export.json
"""
import argparse
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import gpucadforam_b200 as g  # noqa: E402
from gpucadforam_b200 import _capi, synth  # noqa: E402
import ref_py as ref  # noqa: E402

ISO_MASK, BAND_LO, BAND_HI = 0.25, 0.20, 0.30


def digest(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 24), b""):
            h.update(blk)
    return h.hexdigest(), os.path.getsize(path)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fine", type=int, default=512)
    ap.add_argument("--dir", default="/tmp")
    ap.add_argument("--host-verts", type=int, default=30_000_000)
    ap.add_argument("--ref-verts", type=int, default=3_000_000)
    args = ap.parse_args()
    F, R = args.fine, 4
    c = F // R
    coef, harm = synth.gyroid_coefficients(), synth.HARMONICS
    phi = synth.phase_grids(c, c, c, device="cuda", z0=0, cz_total=c, harmonics=harm, periods=F / 40.0)
    ctx = g.Context(0, options=0)
    svl = torch.empty(F ** 3, device="cuda")
    d = (1.0 / R,) * 3
    # count, allocate, extract
    mm = torch.zeros(2, device="cuda")
    g.svl_field(ctx, svl, phi, coef, (c, c, c), (F, F, F), d, d_minmax=mm)
    a, b = [float(x) for x in mm.cpu()]
    act, tot = g.extract_band_raw(ctx, svl, a, b, ISO_MASK, BAND_LO, BAND_HI, (F, F, F), d, (0, 0, 0), None, None, 0, count_only=True)
    mesh = g.MeshBuffers(tot + 3)
    act, tot = g.extract_band_raw(ctx, svl, a, b, ISO_MASK, BAND_LO, BAND_HI, (F, F, F), d, (0, 0, 0), mesh.pos, mesh.norm, tot + 3)
    del svl, phi
    torch.cuda.empty_cache()
    fo = g.File_output(ctx)
    out = {"mesh": "config 3 lattice, %d^3" % F, "total_vertices": tot, "runs": []}

    def run(kind, nverts):
        nverts = min(nverts, tot) // 3 * 3
        path = os.path.join(args.dir, "objbench_%s.obj" % kind)
        torch.cuda.synchronize()
        t0 = time.time()
        if kind == "device":
            fo.file_write_obj(mesh.pos, nverts, path)
        elif kind == "host":
            ctx.set_options(_capi.GCB_OPT_OBJ_HOST)
            fo.file_write_obj(mesh.pos, nverts, path)
            ctx.set_options(0)
        else:
            ref.write_obj(mesh.pos, nverts, path)
        dt = time.time() - t0
        sha, size = digest(path)
        os.remove(path)
        r = {"writer": kind, "vertices": nverts, "seconds": dt, "vertices_per_s": nverts / dt, "file_bytes": size, "sha256": sha}
        out["runs"].append(r)
        return r

    full = run("device", tot)
    hv = run("host", args.host_verts)
    dv = run("device", hv["vertices"])
    assert dv["sha256"] == hv["sha256"], "device and host writers disagree"
    if ref.available():
        rv = run("reference", args.ref_verts)
        dr = run("device", rv["vertices"])
        assert dr["sha256"] == rv["sha256"], "device and reference writers disagree"
        out["speedup_vs_reference_writer_at_%d_vertices" % rv["vertices"]] = rv["seconds"] / dr["seconds"]
    out["speedup_vs_host_writer_at_%d_vertices" % hv["vertices"]] = hv["seconds"] / dv["seconds"]
    out["full_mesh_seconds"] = full["seconds"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
