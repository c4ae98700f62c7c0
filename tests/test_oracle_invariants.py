"""Known-answer style checks of the CPU oracle: the reference ships no tests, so
these are closed-form invariants; the golden-vector tests (test_oracle_golden.py) pin it to the
reference's own CUDA output."""
import numpy as np

import cases
import oracle_py as orc


def _band_case(n=24, typ=0):
    f = orc.normalise_buffer(orc.create_lattice(n, n, n, typ))
    mask, k = orc.normalise_four(f, cases.BAND_LO, cases.BAND_HI)
    return mask, k


def test_scans_and_counts_are_consistent():
    mask, k = _band_case()
    n = mask.shape[0]
    r = orc.extract(orc.MODE_LATTICE_ONE, (n, n, n), (1, 1, 1), (0, 0, 0), cases.ISO_MASK, f0=mask, f1=k, iso1=cases.BAND_LO, iso2=cases.BAND_HI)
    assert r["total"] == int(r["voxelVerts"].sum()) and r["total"] % 3 == 0 and r["total"] > 0
    assert r["active"] == int(r["voxelOccupied"].sum()) == len(r["compVoxelArray"])
    assert np.all(np.diff(r["compVoxelArray"].astype(np.int64)) > 0)
    assert np.array_equal(r["voxelVertsScan"], np.concatenate([[0], np.cumsum(r["voxelVerts"])[:-1]]).astype(np.uint32))
    assert np.array_equal(r["voxelOccupiedScan"], np.concatenate([[0], np.cumsum(r["voxelOccupied"])[:-1]]).astype(np.uint32))
    a, t = orc.count(orc.MODE_LATTICE_ONE, (n, n, n), cases.ISO_MASK, f0=mask, f1=k)
    assert (a, t) == (r["active"], r["total"])
    # w components: pos.w = 1, norm.w = 0 for the lattice variants
    assert np.all(r["pos"][:r["total"], 3] == 1.0) and np.all(r["norm"][:r["total"], 3] == 0.0)
    # all vertices inside the domain; forced-closed faces (A-13): nothing outside [0, n-1]
    p = r["pos"][:r["total"], :3]
    assert p.min() >= 0.0 and p.max() <= n - 1


def test_band_vertices_lie_on_band_edges():
    """Each vertex sits on a grid edge; interpolating k there gives iso1 or iso2 unless it was snapped."""
    mask, k = _band_case(20)
    n = 20
    r = orc.extract(orc.MODE_LATTICE_ONE, (n, n, n), (1, 1, 1), (0, 0, 0), cases.ISO_MASK, f0=mask, f1=k, iso1=cases.BAND_LO, iso2=cases.BAND_HI)
    p = r["pos"][:r["total"], :3].astype(np.float64)
    frac = np.abs(p - np.round(p))
    assert np.all((frac > 1e-6).sum(axis=1) <= 1)  # at most one non-integer coordinate


def test_sphere_mesh_is_closed_and_near_the_sphere():
    """B-3: plain CSG (union of nothing and a sphere), vertices within voxel*sqrt(3) of the sphere, Euler characteristic 2."""
    n, d, rad = 36, 0.5, 6.0
    dyn = orc.sphere((n, n, n), (d, d, d), (0, 0, 0), rad, 2.0, False)
    gp = np.zeros(n * n * n, orc.GP_DTYPE)
    c = (n - 1) / 2.0
    r = orc.extract(orc.MODE_CSG, (n, n, n), (d, d, d), (c, c, c), 0.0, f0=dyn, gp=gp, flags=orc.F_UNION, iso1=0.2, iso2=0.3)
    assert r["total"] > 0
    p = r["pos"][:r["total"], :3].astype(np.float64)
    dist = np.linalg.norm(p, axis=1)
    assert np.all(np.abs(dist - rad) <= d * np.sqrt(3))
    assert np.all(r["norm"][:r["total"], 3] == 0.5)
    # weld and compute V - E + F
    key = np.round(p * 4096).astype(np.int64)
    _, inv = np.unique(key, axis=0, return_inverse=True)
    tris = inv.reshape(-1, 3)
    tris = tris[(tris[:, 0] != tris[:, 1]) & (tris[:, 1] != tris[:, 2]) & (tris[:, 0] != tris[:, 2])]
    edges = np.sort(np.concatenate([tris[:, [0, 1]], tris[:, [1, 2]], tris[:, [2, 0]]]), axis=1)
    ue, cnt = np.unique(edges, axis=0, return_counts=True)
    assert np.all(cnt == 2)  # closed 2-manifold
    V, E, Fc = len(np.unique(tris)), len(ue), len(tris)
    assert V - E + Fc == 2


def test_csg_truth_table():
    """B-2: union s|d, diff (!d)&s, intersect s&d on a 2x2x2 grid with one corner toggled."""
    for s in (0, 1):
        for dbit in (0, 1):
            gp = np.zeros(8, orc.GP_DTYPE)
            gp["val"] = 1
            gp["val"][0] = -1 if s else 1
            dyn = np.ones((2, 2, 2), np.float32)
            dyn[0, 0, 0] = -1.0 if dbit else 1.0
            for flag, expect in ((orc.F_UNION, s | dbit), (orc.F_DIFF, (1 - dbit) & s), (orc.F_INTERSECT, s & dbit)):
                a, t = orc.count(orc.MODE_CSG, (2, 2, 2), 0.0, f0=dyn, gp=gp, flags=flag)
                assert (t == 3) == bool(expect), (s, dbit, flag)


def test_minmax_is_clamped_through_zero():
    """Min_reduction_lattice seeds its lanes with {0,0} (Gratings.cu:1443-1468)."""
    assert orc.minmax(np.array([2.0, 3.0, 5.0], np.float32)) == (0.0, 5.0)
    assert orc.minmax(np.array([-2.0, -3.0], np.float32)) == (-3.0, 0.0)


def test_refine_even_points_copy_and_odd_points_average():
    rng = np.random.RandomState(0)
    c = rng.rand(5, 6, 7).astype(np.float32)
    f = orc.refine(c, (14, 12, 10), (0.5, 0.5, 0.5))
    assert np.array_equal(f[::2, ::2, ::2], c)
    assert np.allclose(f[0, 0, 1:-1:2], 0.5 * (c[0, 0, :-1].astype(np.float64) + c[0, 0, 1:]), rtol=0, atol=1e-7)
    assert np.array_equal(f[-1], f[-2])  # last fine plane clamps to the last control plane (A-12)


def test_obj_writer_welds_and_flips(tmp_path):
    pos = np.array([[0, 0, 0, 1], [1, 0, 0, 1], [0, 1, 0, 1],
                    [1, 0, 0, 1], [1, 1, 0, 1], [0, 1, 0, 1],
                    [0, 0, 0, 1], [1, 0, 0, 1], [0, 1, 0, 1],            # duplicate face -> dropped
                    [2, 2, 2, 1], [2, 2, 2.0004, 1], [3, 3, 3, 1]], np.float32)  # degenerate after quantisation
    path = str(tmp_path / "t.obj")
    assert orc.write_obj(pos, 12, path) == 0
    text = open(path).read()
    assert text.startswith("##Sample latttice new Obj \no Solid \n")
    assert text.count("\nv ") + text.startswith("v ") == 6
    faces = [l for l in text.splitlines() if l.startswith(" f ")]
    assert faces == [" f  1 3 2", " f  2 3 4"]


def test_unit_spectrum_matches_fft_and_is_hermitian():
    """Unit-cell spectrum: the direct DFT of the oracle against numpy's FFT with the reference's
    pick rule (index i mod N, order k, j, i; main.cu:3612-3690); tolerance 1e-6 of the largest coefficient."""
    n, rng = 61, 2
    f = orc.create_lattice(n, n, n, 0)
    got = orc.unit_spectrum(f, rng)
    F = np.fft.fftn(f.astype(np.float64)) / f.size          # axes (z, y, x)
    want = np.array([F[k % n, j % n, i % n] for k in range(-rng, rng + 1) for j in range(-rng, rng + 1) for i in range(-rng, rng + 1)])
    scale = np.abs(want).max()
    assert scale > 1e-3
    assert np.abs(got - want).max() <= 1e-6 * scale
    # real input: c(-k) = conj(c(k)); entry e <-> 124 - e
    assert np.abs(got[::-1] - np.conj(got)).max() <= 1e-6 * scale
    # anisotropic cell exercises the per-axis moduli
    f2 = orc.create_lattice(20, 14, 9, 1)
    g2 = orc.unit_spectrum(f2, 1)
    F2 = np.fft.fftn(f2.astype(np.float64)) / f2.size
    w2 = np.array([F2[k % 9, j % 14, i % 20] for k in (-1, 0, 1) for j in (-1, 0, 1) for i in (-1, 0, 1)])
    assert np.abs(g2 - w2).max() <= 1e-6 * np.abs(w2).max()


def test_mask_helpers_known_answers():
    """primitive_field / topo_field / patch_topo_field against hand-written expectations."""
    n = 64
    gp = np.zeros(n, orc.GP_DTYPE)
    gp["val"] = np.tile(np.array([-1, 0, 1, -1], np.int32), n // 4)
    iso = np.arange(n, dtype=np.float32) - 10
    fmax = np.finfo(np.float32).max
    o = orc.primitive_field(gp, None, iso, True, False)
    assert np.array_equal(o, np.where(gp["val"] > -1, fmax, iso))
    act = np.linspace(-1, 1, n).astype(np.float32)
    act[5] = -0.0
    o = orc.primitive_field(gp, act, iso, False, True)
    assert np.array_equal(o, np.where(act >= 0, fmax, iso)) and o[5] == fmax   # -0.0 >= 0
    o = orc.topo_field(act, iso, 0.25)
    assert np.array_equal(o, np.where(act < np.float32(0.25), np.float32(0), iso))
    # patch_topo_field: the reference guard looks at the layer number only (Isosurface.cu:684-692)
    dims = (4, 2, 8)
    n = dims[0] * dims[1] * dims[2]
    gp = np.zeros(n, orc.GP_DTYPE)
    gp["val"] = 1
    d = np.ones(n, np.float32)
    o = orc.patch_topo_field(d, dims, gp)
    layer = np.arange(n) // (dims[0] * dims[1])
    assert np.array_equal(o == 0, layer < dims[0])   # layers z >= Nx are never patched


def test_lattice_ids_two_zero_use_the_second_field():
    """vertexInterp3_new second half (MarchingCubes_kernel.cu:3347-3413): an edge joining mask ids {2,0} is interpolated on vol_two
    with iso1 / iso2.  One cell column with a single inside corner plane gives closed-form vertices."""
    dims = (2, 2, 2)
    mask = np.zeros(dims[::-1], np.float32)
    mask[1] = 2.0                     # z = 1 plane: id 2 (outside: 2 >= iso), z = 0 plane: id 0 (inside)
    k1 = np.zeros_like(mask)
    k2 = np.zeros_like(mask)
    k2[1] = 0.8                       # crossing of iso1b = 0.6 at t = 0.75 along z
    r = orc.extract(orc.MODE_LATTICE, dims, (1, 1, 1), (0, 0, 0), cases.ISO_MASK, f0=mask, f1=k1, f2=k2, iso1=0.2, iso2=0.3, iso1b=0.6, iso2b=0.9)
    assert r["active"] == 1 and r["total"] == 6
    assert np.allclose(r["pos"][:6, 2], 0.75)


def test_period_and_normalise_three_known_answers():
    """period_data = distance from the shifted axis; GPU_buffer_normalise_three maps it onto [a1, a1 + b1] (main.cu:3929-3931)."""
    dims, d, mean = (16, 12, 8), (1.0, 1.0, 1.0), (8.0, 6.0, 4.0)
    p = orc.period_data(dims, d, mean, "z")
    z, y, x = np.meshgrid(np.arange(8), np.arange(12), np.arange(16), indexing="ij")
    want = np.sqrt((x - 8.0 + 1.0) ** 2 + (y - 6.0 + 1.0) ** 2)
    assert np.allclose(p, want, rtol=1e-6, atol=1e-6)
    assert np.array_equal(p[0], p[5])                       # 'z': no dependence on z
    n3 = orc.normalise_three(p, 1.0, 4.0)                   # NumX/10, NumX/4 in integer arithmetic for NumX = 16
    assert abs(float(n3.min()) - 1.0) < 1e-6 and abs(float(n3.max()) - 5.0) < 1e-6
    th = orc.period_data(dims, d, mean, "z", angle=True)
    assert np.allclose(th, np.arctan2(y - 6.0 + 1.0, x - 8.0 + 1.0), atol=1e-6)
