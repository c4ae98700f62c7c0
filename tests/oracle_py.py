"""ctypes wrapper of oracle/liboracle.so (numpy in, numpy out).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None

MODE_LATTICE_ONE, MODE_LATTICE, MODE_CSG, MODE_TOPO, MODE_REGION = 0, 1, 2, 3, 5
F_UNION, F_DIFF, F_INTERSECT, F_FIXED, F_DYNAMIC, F_MAKE_REGION, F_DISP, F_SHOW_REGION, F_SHOW_DOMAIN = 1, 2, 4, 8, 16, 32, 64, 128, 256

GP_DTYPE = np.dtype([("val", np.int32), ("t_x", np.float32), ("t_y", np.float32), ("t_z", np.float32)])
# triangle_metadata, MarchingCubes_kernel.h:20-32 (64 bytes)
META_DTYPE = np.dtype([("index", np.uint32), ("voxel", np.uint32), ("l_index", np.uint32), ("edge_1", np.uint32), ("edge_2", np.uint32),
                       ("edge_3", np.uint32), ("load_group", np.uint32), ("centroid", np.float32, 3), ("normal", np.float32, 3),
                       ("force_dir", np.float32, 3)])
assert META_DTYPE.itemsize == 64


class McParams(C.Structure):
    _fields_ = [("mode", C.c_int32), ("nx", C.c_uint32), ("ny", C.c_uint32), ("nz", C.c_uint32), ("voxel", C.c_float * 3),
                ("center", C.c_float * 3), ("iso", C.c_float), ("iso1", C.c_float), ("iso2", C.c_float), ("iso1b", C.c_float),
                ("iso2b", C.c_float), ("flags", C.c_uint32), ("max_verts", C.c_uint32), ("f0", C.c_void_p), ("f1", C.c_void_p),
                ("f2", C.c_void_p), ("gp", C.c_void_p), ("disp", C.c_void_p), ("gp2", C.c_void_p), ("meta", C.c_void_p)]


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _v3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def tables():
    tri = np.zeros(256 * 16, np.uint32)
    nv = np.zeros(256, np.uint32)
    lib().orc_tables(_p(tri), _p(nv))
    return tri.reshape(256, 16), nv


def extract(mode, dims, voxel, center, iso, f0=None, f1=None, f2=None, gp=None, disp=None, iso1=0.0, iso2=0.0, iso1b=0.0, iso2b=0.0,
            flags=0, max_verts=None, stages=True, gp2=None, meta=None):
    """dims = (nx, ny, nz) points.  Returns dict with stage arrays, pos, norm, active, total."""
    nx, ny, nz = dims
    ncell = max((nx - 1) * (ny - 1) * (nz - 1), 1)
    if max_verts is None:
        max_verts = max(4 * nx * ny * nz, 300000)
    keep = [None if a is None else _f(a) for a in (f0, f1, f2)]
    gpc = None if gp is None else np.ascontiguousarray(gp)
    dispc = None if disp is None else _f(disp)
    gp2c = None if gp2 is None else np.ascontiguousarray(gp2)
    p = McParams(mode, nx, ny, nz, _v3(voxel), _v3(center), iso, iso1, iso2, iso1b, iso2b, flags, max_verts, _p(keep[0]), _p(keep[1]), _p(keep[2]),
                 _p(gpc), _p(dispc), _p(gp2c), _p(meta))  # meta: caller-allocated META_DTYPE array (REGION + SHOW_REGION)
    out = {}
    names = ["voxelVerts", "voxelOccupied", "voxelVertsScan", "voxelOccupiedScan", "compVoxelArray"]
    arrs = [np.zeros(ncell, np.uint32) if stages else None for _ in names]
    pos = np.zeros((max_verts, 4), np.float32)
    norm = np.zeros((max_verts, 4), np.float32)
    act, tot = C.c_uint32(0), C.c_uint32(0)
    rc = lib().orc_extract(C.byref(p), *[_p(a) for a in arrs], _p(pos), _p(norm), C.byref(act), C.byref(tot))
    assert rc == 0
    for n, a in zip(names, arrs):
        out[n] = a
    if stages:
        out["compVoxelArray"] = out["compVoxelArray"][:act.value]
    out.update(pos=pos, norm=norm, active=act.value, total=tot.value)
    return out


def count(mode, dims, iso, f0=None, f1=None, gp=None, iso1=0.0, iso2=0.0, flags=0):
    keep = [None if a is None else _f(a) for a in (f0, f1)]
    gpc = None if gp is None else np.ascontiguousarray(gp)
    p = McParams(mode, dims[0], dims[1], dims[2], _v3((1, 1, 1)), _v3((0, 0, 0)), iso, iso1, iso2, 0, 0, flags, 0, _p(keep[0]), _p(keep[1]), None,
                 _p(gpc), None)
    a, t = C.c_uint64(0), C.c_uint64(0)
    lib().orc_count(C.byref(p), C.byref(a), C.byref(t))
    return a.value, t.value


def create_lattice(nx, ny, nz, typ):
    out = np.zeros((nz, ny, nx), np.float32)
    lib().orc_create_lattice(_p(out), C.c_uint32(nx), C.c_uint32(ny), C.c_uint32(nz), C.c_uint32(typ))
    return out


def finding_phi(period, dims, ijk, d, latticetype="r", uniform_type=2, const_period=8.0, periods=(8.0, 8.0, 8.0), lcon=0.5, lcon_1=0.05, sinewave_zaxis=False):
    nx, ny, nz = dims
    out = np.zeros(nx * ny * nz, np.float32)
    per = None if period is None else np.ascontiguousarray(period, np.float32)
    lib().orc_finding_phi(_p(out), _p(per), nx, ny, nz, int(ijk[0]), int(ijk[1]), int(ijk[2]), C.c_float(d[0]), C.c_float(d[1]), C.c_float(d[2]), ord(latticetype),
                          uniform_type, C.c_float(const_period), C.c_float(periods[0]), C.c_float(periods[1]), C.c_float(periods[2]), C.c_float(lcon),
                          C.c_float(lcon_1), int(sinewave_zaxis))
    return out


def cg(rhs, dims, iters=500, end_res=0.01):
    """returns (solution, FinalIter, FinalRes)"""
    x = np.ascontiguousarray(rhs, np.float32).copy()
    fi, fr = C.c_int(0), C.c_float(0)
    lib().orc_cg(_p(x), dims[0], dims[1], dims[2], iters, C.c_float(end_res), C.byref(fi), C.byref(fr))
    return x, fi.value, fr.value


def unit_spectrum(f, rng=2):
    """f: [nz, ny, nx] float32 -> complex64 [(2*rng+1)^3] in the reference's lattice_data order."""
    nz, ny, nx = f.shape
    out = np.zeros(((2 * rng + 1) ** 3, 2), np.float32)
    lib().orc_unit_spectrum(_p(np.ascontiguousarray(f, np.float32)), nx, ny, nz, rng, _p(out))
    return out[:, 0] + 1j * out[:, 1]


def _prim(fn, dims, d, *args):
    nx, ny, nz = dims
    out = np.zeros((nz, ny, nx), np.float32)
    getattr(lib(), fn)(_p(out), *args, nx, ny, nz, C.c_float(d[0]), C.c_float(d[1]), C.c_float(d[2]))
    return out


def sphere(dims, d, center, radius, thickness, shell):
    nx, ny, nz = dims
    out = np.zeros((nz, ny, nx), np.float32)
    lib().orc_sphere(_p(out), _v3(center), C.c_float(radius), C.c_float(thickness), nx, ny, nz, C.c_float(d[0]), C.c_float(d[1]), C.c_float(d[2]),
                     int(shell))
    return out


def distance_from_line(dims, d, center, axis, radius, tr, ta, disc):
    nx, ny, nz = dims
    out = np.zeros((nz, ny, nx), np.float32)
    lib().orc_distance_from_line(_p(out), _v3(center), _v3(axis), C.c_float(radius), C.c_float(tr), C.c_float(ta), nx, ny, nz, C.c_float(d[0]),
                                 C.c_float(d[1]), C.c_float(d[2]), int(disc))
    return out


def cuboid(dims, d, center, angles, xw, yw, zw):
    return _prim("orc_cuboid", dims, d, _v3(center), _v3(angles), C.c_float(xw), C.c_float(yw), C.c_float(zw))


def cuboid_shell(dims, d, center, angles, xw, yw, zw, th):
    return _prim("orc_cuboid_shell", dims, d, _v3(center), _v3(angles), C.c_float(xw), C.c_float(yw), C.c_float(zw), C.c_float(th))


def torus(dims, d, center, angles, R, rc):
    return _prim("orc_torus", dims, d, _v3(center), _v3(angles), C.c_float(R), C.c_float(rc))


def cone(dims, d, center, angles, br, h):
    return _prim("orc_cone", dims, d, _v3(center), _v3(angles), C.c_float(br), C.c_float(h))


def cone_frustum(dims, d, center, angles, tr, br, h):
    return _prim("orc_cone_frustum", dims, d, _v3(center), _v3(angles), C.c_float(tr), C.c_float(br), C.c_float(h))


def pyramid_frustum(dims, d, center, angles, xb, xt, yh, zb, zt):
    return _prim("orc_pyramid_frustum", dims, d, _v3(center), _v3(angles), C.c_float(xb), C.c_float(xt), C.c_float(yh), C.c_float(zb), C.c_float(zt))


def minmax(f):
    f = _f(f)
    a, b = C.c_float(0), C.c_float(0)
    lib().orc_minmax(_p(f), C.c_size_t(f.size), C.byref(a), C.byref(b))
    return a.value, b.value


def normalise_buffer(f):
    f = _f(f)
    out = np.zeros_like(f)
    lib().orc_normalise_buffer(_p(f), _p(out), C.c_size_t(f.size))
    return out


def normalise_four(f, iso1, iso2, ab=None):
    f = _f(f)
    nz, ny, nx = f.shape
    mask, k = np.zeros_like(f), np.zeros_like(f)
    if ab is None:
        lib().orc_normalise_four(_p(f), _p(mask), _p(k), nx, ny, nz, C.c_float(iso1), C.c_float(iso2))
    else:
        lib().orc_normalise_four_ab(_p(f), _p(mask), _p(k), nx, ny, nz, C.c_float(iso1), C.c_float(iso2), C.c_float(ab[0]), C.c_float(ab[1]))
    return mask, k


def refine(coarse, fdims, d):
    coarse = _f(coarse)
    cz, cy, cx = coarse.shape
    out = np.zeros((fdims[2], fdims[1], fdims[0]), np.float32)
    lib().orc_refine(_p(coarse), cx, cy, cz, _p(out), fdims[0], fdims[1], fdims[2], C.c_float(d[0]), C.c_float(d[1]), C.c_float(d[2]))
    return out


def svl_field(phi, coef, fdims, d):
    """phi [nh, cz, cy, cx]; returns the accumulated field [nz2, ny2, nx2]."""
    phi = _f(phi)
    nh, cz, cy, cx = phi.shape
    svl = np.zeros((fdims[2], fdims[1], fdims[0]), np.float32)
    for h in range(nh):
        lib().orc_svl_accumulate(_p(svl), _p(phi[h]), cx, cy, cz, fdims[0], fdims[1], fdims[2], C.c_float(d[0]), C.c_float(d[1]), C.c_float(d[2]),
                                 C.c_float(coef[h][0]), C.c_float(coef[h][1]))
    return svl


def copy_parameter(vol_one, vol_two, vol_lattice, dims, iso, dynamic=False, iso1=0.2, iso2=0.3, obj_union=True, obj_diff=False, obj_intersect=False):
    nx, ny, nz = dims
    v2 = None if vol_two is None else _f(vol_two)
    vl = None if vol_lattice is None else _f(vol_lattice)
    lib().orc_copy_parameter(_p(vol_one), _p(v2), _p(vl), int(dynamic), C.c_float(iso1), C.c_float(iso2), nx, ny, nz, C.c_float(iso), int(obj_union),
                             int(obj_diff), int(obj_intersect))
    return vol_one


def primitive_field(prim, active, isosurf, fixed, dynamic):
    """orc_primitive_field (Gratings.cu:1695-1725): returns the patched copy of isosurf."""
    out = _f(isosurf).copy()
    a = None if active is None else _f(active)
    lib().orc_primitive_field(_p(prim), _p(a), _p(out), C.c_size_t(out.size), int(fixed), int(dynamic))
    return out


def topo_field(topo, isosurf, volfrac):
    """orc_topo_field (Gratings.cu:1666-1681)."""
    out = _f(isosurf).copy()
    t = _f(topo)
    lib().orc_topo_field(_p(t), _p(out), C.c_float(volfrac), C.c_size_t(out.size))
    return out


def patch_topo_field(d, dims, vol_one):
    """orc_patch_topo_field (Isosurface.cu:674-707), index guard restated literally."""
    out = _f(d).copy()
    lib().orc_patch_topo_field(_p(out), dims[0], dims[1], dims[2], _p(vol_one))
    return out


def period_data(dims, d, mean, axis="z", angle=False):
    """orc_period_data / orc_angle_data (Gratings.cu:775-853) -> [nz, ny, nx] float32."""
    nx, ny, nz = dims
    out = np.zeros((nz, ny, nx), np.float32)
    fn = lib().orc_angle_data if angle else lib().orc_period_data
    fn(_p(out), nx, ny, nz, C.c_float(d[0]), C.c_float(d[1]), C.c_float(d[2]), C.c_float(mean[0]), C.c_float(mean[1]), C.c_float(mean[2]), ord(axis))
    return out


def normalise_three(f, a1, b1):
    """orc_normalise_three (Gratings.cu:1539-1572)."""
    f = _f(f)
    out = np.zeros_like(f)
    lib().orc_normalise_three(_p(f), _p(out), C.c_size_t(f.size), C.c_float(a1), C.c_float(b1))
    return out


def write_obj(pos, total_verts, filename):
    pos = _f(pos)
    return lib().orc_write_obj(_p(pos), C.c_uint32(total_verts), filename.encode())


def num_threads():
    return lib().orc_num_threads()
