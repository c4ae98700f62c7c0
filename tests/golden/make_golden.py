#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the REFERENCE's own CUDA kernels (oracle/_ref/libgpucad_ref.so,
built by oracle/Makefile from the unmodified sources under /root/reference/src) on a GPU.

The reference ships no tests or fixtures, so these files are the golden vectors that pin the CPU oracle
(tests/test_oracle_golden.py, runs without a GPU).  Run on the GPU box:
    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'
then copy gpurun_out/golden/*.npz to tests/golden/ and commit them.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cases  # noqa: E402
import ref_py as ref  # noqa: E402
from gpu_util import dev, gp_zeros, gp_to_numpy, max_verts_for, stage_dict  # noqa: E402
import gpucadforam_b200 as g  # noqa: E402  (only for Scratch/MeshBuffers containers: plain torch allocations)


def pack(scr, mesh, dims, act, tot):
    ncell = (dims[0] - 1) * (dims[1] - 1) * (dims[2] - 1)
    d = stage_dict(scr, ncell, act)
    d.update(pos=mesh.pos[:tot].cpu().numpy(), norm=mesh.norm[:tot].cpu().numpy(), active=np.uint32(act), total=np.uint32(tot))
    return d


def region(out):
    """computeIsosurface_region (f-4): the three display modes on retained grids, incl. the triangle_metadata records."""
    R = cases.REGION
    dims, d = R["dims"], R["d"]
    npts = dims[0] * dims[1] * dims[2]
    zero = torch.zeros(npts, device="cuda")
    f = torch.zeros(npts, device="cuda")
    vol_topo, vol_one = gp_zeros(npts), gp_zeros(npts)
    s, c, y = R["topo_sphere"], R["cuboid"], R["dyn_sphere"]
    ref.sphere(f, s["center"], s["radius"], s["thickness"], dims, d, False)
    ref.copy_parameter(vol_topo, f, zero, dims, d, 0.0, obj_union=True)
    ref.cuboid(f, c["center"], c["angles"], c["xw"], c["yw"], c["zw"], dims, d)
    ref.copy_parameter(vol_one, f, zero, dims, d, 0.0, obj_union=True)
    dyn = torch.zeros(npts, device="cuda")
    ref.sphere(dyn, y["center"], y["radius"], y["thickness"], dims, d, False)
    mv = max_verts_for(dims)
    ncell = (dims[0] - 1) * (dims[1] - 1) * (dims[2] - 1)
    res = dict(vol_topo=vol_topo.cpu().numpy(), vol_one=vol_one.cpu().numpy(), dynamic=dyn.cpu().numpy())
    for mode in cases.REGION_MODES:
        scr, mesh = g.Scratch(ncell), g.MeshBuffers(mv)
        meta = torch.full((mv // 3, 16), 0x7f7f7f7f, dtype=torch.int32, device="cuda")
        a, t = ref.isosurface_region(False, mesh.pos, mesh.norm, 0.0, dims, d, (0, 0, 0), scr, mv, vol_topo, vol_one, dyn, triangle_data=meta,
                                     **{mode: True})
        for k, v in pack(scr, mesh, dims, a, t).items():
            res[mode + "_" + k] = v
        res[mode + "_meta"] = meta[:t // 3].cpu().numpy()
        print("region", mode, a, t)
    np.savez_compressed(os.path.join(out, "region.npz"), **res)


def main(out, only=None):
    os.makedirs(out, exist_ok=True)
    if only == "region":
        return region(out)
    region(out)
    # ---- gyroid unit cell, band extraction (config 1 in miniature)
    n = cases.GYROID["n"]
    raw = torch.zeros(n ** 3, device="cuda")
    ref.create_lattice(raw, n, n, n, cases.GYROID["type"])
    f = torch.zeros_like(raw)
    ref.normalise_buffer(raw, f, n ** 3)
    mask, k = torch.zeros_like(f), torch.zeros_like(f)
    ref.normalise_four(f, mask, k, (n, n, n), cases.BAND_LO, cases.BAND_HI)
    dims = (n, n, n)
    mv = max_verts_for(dims)
    scr, mesh = g.Scratch((n - 1) ** 3), g.MeshBuffers(mv)
    a, t = ref.isosurface_lattice(True, False, mask, mesh.pos, mesh.norm, cases.ISO_MASK, dims, (1, 1, 1), (0, 0, 0), scr, mv, k, None, cases.BAND_LO,
                                  cases.BAND_HI)
    np.savez_compressed(os.path.join(out, "gyroid_band.npz"), raw=raw.cpu().numpy(), normalised=f.cpu().numpy(), mask=mask.cpu().numpy(), k=k.cpu().numpy(),
                        **pack(scr, mesh, dims, a, t))
    print("gyroid_band", a, t)
    tp = {}
    for typ in cases.TPMS_TYPES:
        o = torch.zeros(17 ** 3, device="cuda")
        ref.create_lattice(o, 17, 17, 17, typ)
        tp["type%d" % typ] = o.cpu().numpy()
    np.savez_compressed(os.path.join(out, "tpms_types.npz"), **tp)

    # ---- CSG: sphere U box - cylinder (config 2 in miniature)
    C = cases.CSG
    dims, d = C["dims"], C["d"]
    npts = dims[0] * dims[1] * dims[2]
    vol_one = gp_zeros(npts)
    boundary = torch.zeros(npts, device="cuda")
    lattice = torch.zeros(npts, device="cuda")  # d_volumethree, read unconditionally by classify_copy_Voxel
    s, c, y = C["sphere"], C["cuboid"], C["cylinder"]
    ref.sphere(boundary, s["center"], s["radius"], s["thickness"], dims, d, False)
    sphere_f = boundary.cpu().numpy().copy()
    ref.copy_parameter(vol_one, boundary, lattice, dims, d, 0.0, obj_union=True)
    ref.cuboid(boundary, c["center"], c["angles"], c["xw"], c["yw"], c["zw"], dims, d)
    cuboid_f = boundary.cpu().numpy().copy()
    ref.copy_parameter(vol_one, boundary, lattice, dims, d, 0.0, obj_union=True)
    ref.distance_from_line(boundary, y["center"], y["axis"], y["radius"], y["tr"], y["ta"], dims, d, False)
    mv = max_verts_for(dims)
    ncell = (dims[0] - 1) * (dims[1] - 1) * (dims[2] - 1)
    scr, mesh = g.Scratch(ncell), g.MeshBuffers(mv)
    a, t = ref.isosurface_csg(False, mesh.pos, mesh.norm, 0.0, dims, d, (0, 0, 0), scr, mv, vol_one, boundary, None, obj_union=False, obj_diff=True)
    ref.write_obj(mesh.pos, t, os.path.join(out, "csg_ref.obj"))
    np.savez_compressed(os.path.join(out, "csg.npz"), sphere=sphere_f, cuboid=cuboid_f, cylinder=boundary.cpu().numpy(),
                        vol_one=vol_one.cpu().numpy(), obj=np.frombuffer(open(os.path.join(out, "csg_ref.obj"), "rb").read(), np.uint8),
                        **pack(scr, mesh, dims, a, t))
    os.remove(os.path.join(out, "csg_ref.obj"))
    print("csg", a, t)

    # ---- SVL field (configs 3/4 in miniature), ratio 2 and ratio 4, then band extraction
    for name, cfg in (("svl", cases.SVL), ("svl4", cases.SVL4)):
        phi, coef = cases.svl_inputs(cfg)
        cx, cy, cz = cfg["cdims"]
        fx, fy, fz = cfg["fdims"]
        ref.setup_texture(cx, cy, cz)
        svl = torch.zeros(fx * fy * fz, device="cuda")
        ga = torch.zeros((fx * fy * fz, 2), device="cuda")
        ref.svl_field(svl, ga, dev(phi), len(coef), dev(np.array(coef, np.float32)), cfg["cdims"], cfg["fdims"], cfg["d"])
        up = torch.zeros_like(svl)
        ref.upload_texture(dev(phi[0]), cx, cy, cz)
        ref.refine(up, cfg["fdims"], cfg["d"])
        ref.delete_texture()
        mask, k = torch.zeros_like(svl), torch.zeros_like(svl)
        ref.normalise_four(svl, mask, k, cfg["fdims"], cases.BAND_LO, cases.BAND_HI)
        dims = cfg["fdims"]
        mv = max_verts_for(dims)
        scr, mesh = g.Scratch((fx - 1) * (fy - 1) * (fz - 1)), g.MeshBuffers(mv)
        a, t = ref.isosurface_lattice(False, False, mask, mesh.pos, mesh.norm, cases.ISO_MASK, dims, cfg["d"], (0, 0, 0), scr, mv, k, torch.zeros_like(k),
                                      cases.BAND_LO, cases.BAND_HI, 0.0, 0.0)
        np.savez_compressed(os.path.join(out, name + ".npz"), phi=phi, coef=np.array(coef, np.float32), svl=svl.cpu().numpy(), refined0=up.cpu().numpy(),
                            mask=mask.cpu().numpy(), k=k.cpu().numpy(), **pack(scr, mesh, dims, a, t))
        print(name, a, t)

    # ---- topology-optimised density (config 5 in miniature)
    T = cases.TOPO
    coarse = cases.topo_coarse(T)
    cx, cy, cz = T["cdims"]
    fx, fy, fz = T["fdims"]
    ref.setup_texture(cx, cy, cz)
    ref.upload_texture(dev(coarse), cx, cy, cz)
    dens = torch.zeros(fx * fy * fz, device="cuda")
    ref.refine(dens, T["fdims"], T["d"])
    ref.delete_texture()
    npts = fx * fy * fz
    rng = np.random.RandomState(3)
    host = np.zeros(npts, dtype=[("val", np.int32), ("t_x", np.float32), ("t_y", np.float32), ("t_z", np.float32)])
    host["t_y"] = np.where(rng.rand(npts) < 0.2, rng.rand(npts), 0).astype(np.float32)
    result = rng.rand(npts).astype(np.float32)
    vol_topo = torch.from_numpy(host.view(np.int32).reshape(-1, 4)).cuda()
    dims = T["fdims"]
    mv = max_verts_for(dims)
    scr, mesh = g.Scratch((fx - 1) * (fy - 1) * (fz - 1)), g.MeshBuffers(mv)
    a, t = ref.isosurface_topo(False, mesh.pos, mesh.norm, T["iso"], dims, T["d"], (0, 0, 0), scr, mv, vol_topo, dens, 0.0, dev(result), vol_one=vol_topo,
                               d_solid=dens)
    np.savez_compressed(os.path.join(out, "topo.npz"), coarse=coarse, density=dens.cpu().numpy(), vol_topo=host.view(np.int32).reshape(-1, 4), result=result,
                        **pack(scr, mesh, dims, a, t))
    print("topo", a, t)

    # ---- SVL phase solve (f-2): right-hand sides and CG solutions of the reference kernels
    P = cases.PHASE
    per = cases.phase_period(P)
    dper = dev(per)
    npts = P["dims"][0] * P["dims"][1] * P["dims"][2]
    ph = {}
    for lt, ut in (("r", 2), ("b", 0), ("n", 1), ("s", 2)):
        for hi, h in enumerate(P["harmonics"]):
            b = torch.zeros(npts, device="cuda")
            ref.finding_phi(b, dper, P["dims"], h, P["d"], latticetype=lt, uniform_type=ut, const_period=7.3, periods=(6.1, 7.7, 5.3), lcon=0.45, lcon_1=0.07,
                            sinewave_zaxis=(lt == "s"))
            ph["rhs_%s%d_%d" % (lt, ut, hi)] = b.cpu().numpy()
            if lt == "r":
                x = b.clone()
                fi, fr = ref.cg(x, P["dims"], P["iters"], P["end_res"])
                ph["sol_%d" % hi] = x.cpu().numpy()
                ph["iters_%d" % hi] = np.int32(fi)
                ph["res_%d" % hi] = np.float32(fr)
    np.savez_compressed(os.path.join(out, "phase_solve.npz"), period=per, **ph)
    print("phase_solve", [int(ph["iters_%d" % i]) for i in range(len(P["harmonics"]))])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE), sys.argv[2] if len(sys.argv) > 2 else None)
