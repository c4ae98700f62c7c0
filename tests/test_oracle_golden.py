"""Pin the CPU oracle to the REFERENCE's own output: tests/golden/*.npz were produced by the unmodified
reference CUDA kernels on a B200 (tests/golden/make_golden.py).  Runs without a GPU.

Extraction is compared on the reference's own fp32 fields (stage arrays, counts and topology bit-exact;
vertices/normals within 1e-5 relative).  Field producers are compared with an ulp-level tolerance because
glibc and libdevice sinf/cosf differ in the last bit."""
import os

import numpy as np
import pytest

import cases
import oracle_py as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    p = os.path.join(GOLD, name + ".npz")
    if not os.path.exists(p):
        pytest.skip("golden fixture %s not generated yet" % name)
    return np.load(p)


def check_mesh(o, gold, what):
    assert o["active"] == int(gold["active"]) and o["total"] == int(gold["total"]), what
    for k in ("voxelVerts", "voxelOccupied", "voxelVertsScan", "voxelOccupiedScan", "compVoxelArray"):
        assert np.array_equal(o[k], gold[k]), "%s: %s" % (what, k)
    t = o["total"]
    scale = max(1.0, float(np.abs(gold["pos"][:, :3]).max()))
    assert np.allclose(o["pos"][:t], gold["pos"], rtol=0, atol=1e-5 * scale), "%s: pos %g" % (what, np.abs(o["pos"][:t] - gold["pos"]).max())
    n0, n1 = o["norm"][:t, :3].astype(np.float64), gold["norm"][:, :3].astype(np.float64)
    nscale = np.maximum(np.linalg.norm(n1, axis=1, keepdims=True), 1e-3)
    assert np.all(np.abs(n0 - n1) <= 1e-4 * nscale + 1e-7), "%s: normals" % what
    assert np.array_equal(o["norm"][:t, 3], gold["norm"][:, 3]), "%s: norm.w" % what


def test_gyroid_band_extraction_matches_reference():
    gd = load("gyroid_band")
    n = cases.GYROID["n"]
    mask, k = gd["mask"].reshape(n, n, n), gd["k"].reshape(n, n, n)
    o = orc.extract(orc.MODE_LATTICE_ONE, (n, n, n), (1, 1, 1), (0, 0, 0), cases.ISO_MASK, f0=mask, f1=k, iso1=cases.BAND_LO, iso2=cases.BAND_HI)
    check_mesh(o, gd, "gyroid band")


def test_gyroid_field_and_normalisation_match_reference():
    gd = load("gyroid_band")
    n = cases.GYROID["n"]
    raw = orc.create_lattice(n, n, n, cases.GYROID["type"]).reshape(-1)
    assert np.allclose(raw, gd["raw"], rtol=0, atol=4e-6)
    # normalisation of the reference's own raw field is exact arithmetic (sub, sub, IEEE div)
    norm = orc.normalise_buffer(gd["raw"])
    assert np.array_equal(norm, gd["normalised"])
    mask, k = orc.normalise_four(gd["normalised"].reshape(n, n, n), cases.BAND_LO, cases.BAND_HI)
    assert np.array_equal(mask.reshape(-1), gd["mask"]) and np.array_equal(k.reshape(-1), gd["k"])


def test_tpms_types_match_reference():
    gd = load("tpms_types")
    for typ in cases.TPMS_TYPES:
        o = orc.create_lattice(17, 17, 17, typ).reshape(-1)
        assert np.allclose(o, gd["type%d" % typ], rtol=0, atol=6e-6), "TPMS type %d: %g" % (typ, np.abs(o - gd["type%d" % typ]).max())


def test_csg_pipeline_matches_reference(tmp_path):
    gd = load("csg")
    C = cases.CSG
    dims, d = C["dims"], C["d"]
    s, c, y = C["sphere"], C["cuboid"], C["cylinder"]
    # primitive fields (ulp-level tolerance: rotation matrix from glibc vs libdevice trig)
    assert np.allclose(orc.sphere(dims, d, s["center"], s["radius"], s["thickness"], False).reshape(-1), gd["sphere"], rtol=0, atol=1e-4)
    assert np.allclose(orc.cuboid(dims, d, c["center"], c["angles"], c["xw"], c["yw"], c["zw"]).reshape(-1), gd["cuboid"], rtol=0, atol=1e-4)
    assert np.allclose(orc.distance_from_line(dims, d, y["center"], y["axis"], y["radius"], y["tr"], y["ta"], False).reshape(-1), gd["cylinder"], rtol=0,
                       atol=1e-4)
    # retain on the reference's own fields: bit-exact grid_points
    vol_one = np.zeros(dims[0] * dims[1] * dims[2], orc.GP_DTYPE)
    orc.copy_parameter(vol_one, gd["sphere"], None, dims, 0.0, obj_union=True)
    orc.copy_parameter(vol_one, gd["cuboid"], None, dims, 0.0, obj_union=True)
    assert np.array_equal(vol_one.view(np.int32).reshape(-1, 4), gd["vol_one"])
    o = orc.extract(orc.MODE_CSG, dims, d, (0, 0, 0), 0.0, f0=gd["cylinder"], gp=vol_one, flags=orc.F_DIFF, iso1=0.2, iso2=0.3)
    check_mesh(o, gd, "csg")
    # .obj writer on the reference's own vertices: identical bytes
    path = str(tmp_path / "o.obj")
    orc.write_obj(gd["pos"], int(gd["total"]), path)
    assert open(path, "rb").read() == gd["obj"].tobytes()


@pytest.mark.parametrize("name,cfg", [("svl", cases.SVL), ("svl4", cases.SVL4)])
def test_svl_matches_reference(name, cfg):
    gd = load(name)
    phi, coef = gd["phi"], [tuple(c) for c in gd["coef"]]
    fx, fy, fz = cfg["fdims"]
    up = orc.refine(phi[0], cfg["fdims"], cfg["d"]).reshape(-1)
    assert np.allclose(up, gd["refined0"], rtol=0, atol=1e-5), "trilinear upsample vs texture unit: %g" % np.abs(up - gd["refined0"]).max()
    svl = orc.svl_field(phi, coef, cfg["fdims"], cfg["d"]).reshape(-1)
    assert np.allclose(svl, gd["svl"], rtol=0, atol=3e-5), "SVL field: %g" % np.abs(svl - gd["svl"]).max()
    mask, k = orc.normalise_four(gd["svl"].reshape(fz, fy, fx), cases.BAND_LO, cases.BAND_HI)
    assert np.array_equal(mask.reshape(-1), gd["mask"]) and np.array_equal(k.reshape(-1), gd["k"])
    o = orc.extract(orc.MODE_LATTICE, cfg["fdims"], cfg["d"], (0, 0, 0), cases.ISO_MASK, f0=gd["mask"], f1=gd["k"], f2=np.zeros_like(gd["k"]),
                    iso1=cases.BAND_LO, iso2=cases.BAND_HI)
    check_mesh(o, gd, name)


def test_topo_matches_reference():
    gd = load("topo")
    T = cases.TOPO
    dens = orc.refine(gd["coarse"], T["fdims"], T["d"]).reshape(-1)
    assert np.allclose(dens, gd["density"], rtol=0, atol=1e-6)
    gp = np.ascontiguousarray(gd["vol_topo"]).view(orc.GP_DTYPE).reshape(-1)
    o = orc.extract(orc.MODE_TOPO, T["fdims"], T["d"], (0, 0, 0), T["iso"], f0=gd["density"], f1=gd["result"], gp=gp, iso1=0.0)
    check_mesh(o, gd, "topo")


def test_phase_solve_against_reference_kernels():
    """SVL phase solve: right-hand sides of the reference's finding_phi kernel (tolerance: host libm vs
    libdevice atan2f/sinf/cosf) and its CG -- the oracle's CG on the reference's right-hand side must reproduce the reference's
    solution, iteration count and residual bit for bit (only +, *, / and the reduction trees are involved)."""
    G = np.load(os.path.join(GOLD, "phase_solve.npz"))
    P = cases.PHASE
    dims, d = P["dims"], P["d"]
    per = G["period"]
    assert np.array_equal(per, cases.phase_period(P))
    for lt, ut in (("r", 2), ("b", 0), ("n", 1), ("s", 2)):
        for hi, h in enumerate(P["harmonics"]):
            want = G["rhs_%s%d_%d" % (lt, ut, hi)]
            got = orc.finding_phi(per, dims, h, d, latticetype=lt, uniform_type=ut, const_period=7.3, periods=(6.1, 7.7, 5.3), lcon=0.45, lcon_1=0.07,
                                  sinewave_zaxis=(lt == "s"))
            assert np.abs(got - want).max() <= 1e-5 * max(1.0, float(np.abs(want).max())), "finding_phi %s/%d %s" % (lt, ut, h)
    for hi, h in enumerate(P["harmonics"]):
        x, fi, fr = orc.cg(G["rhs_r2_%d" % hi], dims, P["iters"], P["end_res"])
        assert fi == int(G["iters_%d" % hi]) and np.float32(fr) == G["res_%d" % hi]
        assert np.array_equal(x.view(np.uint32), G["sol_%d" % hi].view(np.uint32)), "CG solution of harmonic %s" % (h,)


@pytest.mark.parametrize("mode", cases.REGION_MODES)
def test_region_matches_reference(mode):
    """computeIsosurface_region on the reference's own retained grids: cascade / counts / stage arrays / aa / integer
    metadata bit-exact, vertices 1e-5 relative, unit normals 1e-4 (rsqrtf is MUFU.RSQ on the GPU, 1/sqrtf here)."""
    gd = load("region")
    R = cases.REGION
    dims, d = R["dims"], R["d"]
    gp = np.ascontiguousarray(gd["vol_one"]).view(orc.GP_DTYPE).reshape(-1)
    gp2 = np.ascontiguousarray(gd["vol_topo"]).view(orc.GP_DTYPE).reshape(-1)
    flags = {"make_region": orc.F_MAKE_REGION, "show_region": orc.F_SHOW_REGION, "show_domain": orc.F_SHOW_DOMAIN}[mode]
    tot = int(gd[mode + "_total"])
    meta = np.zeros(max(tot // 3, 1), orc.META_DTYPE)
    meta.view(np.int32)[:] = 0x7f7f7f7f
    o = orc.extract(orc.MODE_REGION, dims, d, (0, 0, 0), 0.0, f0=gd["dynamic"], gp=gp, gp2=gp2, flags=flags, meta=meta)
    assert (o["active"], o["total"]) == (int(gd[mode + "_active"]), tot)
    for k in ("voxelVerts", "voxelOccupied", "voxelVertsScan", "voxelOccupiedScan", "compVoxelArray"):
        assert np.array_equal(o[k], gd[mode + "_" + k]), k
    pos, norm = gd[mode + "_pos"], gd[mode + "_norm"]
    scale = max(1.0, float(np.abs(pos[:, :3]).max()))
    assert np.allclose(o["pos"][:tot], pos, rtol=0, atol=1e-5 * scale)
    assert np.array_equal(o["norm"][:tot, 3], norm[:, 3])
    e1, e2 = (pos[1::3, :3] - pos[0::3, :3]).astype(np.float64), (pos[2::3, :3] - pos[0::3, :3]).astype(np.float64)
    solid = np.repeat(np.linalg.norm(np.cross(e1, e2), axis=1) > 1e-3, 3)
    assert solid.sum() > 0.5 * tot
    assert np.allclose(o["norm"][:tot][solid, :3], norm[solid, :3], rtol=0, atol=1e-4)
    gm = np.ascontiguousarray(gd[mode + "_meta"]).view(orc.META_DTYPE).reshape(-1)
    if mode == "show_region":
        for k in ("index", "voxel", "l_index", "edge_1", "edge_2", "edge_3", "load_group"):
            assert np.array_equal(meta[:tot // 3][k], gm[k]), k
        assert np.allclose(meta[:tot // 3]["centroid"], gm["centroid"], rtol=0, atol=1e-5 * scale)
    else:
        assert (gm.view(np.int32) == 0x7f7f7f7f).all()  # the reference leaves triangle_data alone outside show_region
