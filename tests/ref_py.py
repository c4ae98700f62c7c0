"""ctypes wrapper of oracle/_ref/libgpucad_ref.so: the UNMODIFIED reference CUDA kernels behind the
headless harness oracle/ref_harness.cu.  TEST / BASELINE INFRASTRUCTURE ONLY (needs a GPU)."""
import ctypes as C
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "oracle", "_ref", "libgpucad_ref.so")
_lib = None


def available():
    return os.path.exists(PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(PATH)
        assert _lib.ref_init() == 0
    return _lib


def _p(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous()
    return C.c_void_p(t.data_ptr())


F = C.c_float


def _ok(rc):
    assert rc == 0, "reference harness call failed"


def create_lattice(out, nx, ny, nz, typ):
    _ok(lib().ref_create_lattice(_p(out), C.c_uint(nx), C.c_uint(ny), C.c_uint(nz), C.c_uint(typ)))


def sphere(out, center, r, t, dims, d, shell):
    _ok(lib().ref_sphere(_p(out), F(center[0]), F(center[1]), F(center[2]), F(r), F(t), dims[0], dims[1], dims[2], F(d[0]), F(d[1]), F(d[2]), int(shell)))


def distance_from_line(out, center, axis, r, tr, ta, dims, d, disc):
    _ok(lib().ref_distance_from_line(_p(out), F(center[0]), F(center[1]), F(center[2]), F(axis[0]), F(axis[1]), F(axis[2]), F(r), F(tr), F(ta), dims[0],
                                     dims[1], dims[2], F(d[0]), F(d[1]), F(d[2]), int(disc)))


def _rot(fn, out, center, ang, params, dims, d):
    _ok(getattr(lib(), fn)(_p(out), F(center[0]), F(center[1]), F(center[2]), F(ang[0]), F(ang[1]), F(ang[2]), *[F(p) for p in params], dims[0], dims[1],
                           dims[2], F(d[0]), F(d[1]), F(d[2])))


def cuboid(out, center, ang, xw, yw, zw, dims, d):
    _rot("ref_cuboid", out, center, ang, (xw, yw, zw), dims, d)


def cuboid_shell(out, center, ang, xw, yw, zw, th, dims, d):
    _rot("ref_cuboid_shell", out, center, ang, (xw, yw, zw, th), dims, d)


def torus(out, center, ang, R, rc, dims, d):
    _rot("ref_torus", out, center, ang, (R, rc), dims, d)


def cone(out, center, ang, br, h, dims, d):
    _rot("ref_cone", out, center, ang, (br, h), dims, d)


def cone_frustum(out, center, ang, tr, br, h, dims, d):
    _rot("ref_cone_frustum", out, center, ang, (tr, br, h), dims, d)


def pyramid_frustum(out, center, ang, xb, xt, yh, zb, zt, dims, d):
    _rot("ref_pyramid_frustum", out, center, ang, (xb, xt, yh, zb, zt), dims, d)


def normalise_buffer(inp, out, n):
    _ok(lib().ref_normalise_buffer(_p(inp), _p(out), int(n)))


def normalise_four(inp, mask, k, dims, iso1, iso2):
    _ok(lib().ref_normalise_four(_p(inp), _p(mask), _p(k), C.c_size_t(dims[0] * dims[1] * dims[2]), dims[0], dims[1], dims[2], F(iso1), F(iso2)))


def setup_texture(cx, cy, cz):
    _ok(lib().ref_setup_texture(cx, cy, cz))


def upload_texture(phi, cx, cy, cz):
    _ok(lib().ref_upload_texture(_p(phi), cx, cy, cz))


def delete_texture():
    _ok(lib().ref_delete_texture())


def refine(out, fdims, d):
    _ok(lib().ref_refine(_p(out), fdims[0], fdims[1], fdims[2], F(d[0]), F(d[1]), F(d[2])))


def svl_field(svl, ga, phi, nh, coef_dev, cdims, fdims, d):
    _ok(lib().ref_svl_field(_p(svl), _p(ga), _p(phi), nh, _p(coef_dev), cdims[0], cdims[1], cdims[2], fdims[0], fdims[1], fdims[2], F(d[0]), F(d[1]),
                            F(d[2])))


def _zeros_like_grid(dims, like):
    return torch.zeros(dims[0] * dims[1] * dims[2], device=like.device, dtype=torch.float32)


def copy_parameter(vol_one, vol_two, vol_lattice, dims, voxel, iso, fixed=False, dynamic=False, iso1=0.2, iso2=0.3, obj_union=True, obj_diff=False,
                   obj_intersect=False):
    # classify_copy_Voxel reads BOTH float fields unconditionally (MarchingCubes_kernel.cu:177-178): the app always passes valid buffers
    if vol_lattice is None:
        vol_lattice = _zeros_like_grid(dims, vol_one)
    if vol_two is None:
        vol_two = _zeros_like_grid(dims, vol_one)
    _ok(lib().ref_copy_parameter(None, F(iso), C.c_uint(dims[0]), C.c_uint(dims[1]), C.c_uint(dims[2]), F(voxel[0]), F(voxel[1]), F(voxel[2]),
                                 _p(vol_one), _p(vol_two), _p(vol_lattice), int(fixed), int(dynamic), F(iso1), F(iso2), int(obj_union), int(obj_diff),
                                 int(obj_intersect)))


def primitive_field(prim, active, isosurf, fixed, dynamic, dims):
    _ok(lib().ref_primitive_field(_p(prim), _p(active), _p(isosurf), int(fixed), int(dynamic), dims[0], dims[1], dims[2]))


def topo_field(topo, isosurf, volfrac, dims):
    _ok(lib().ref_topo_field(_p(topo), _p(isosurf), F(volfrac), dims[0], dims[1], dims[2]))


def patch_topo_field(d, dims, vol_one):
    _ok(lib().ref_patch_topo_field(_p(d), dims[0], dims[1], dims[2], _p(vol_one)))


def _scr(s):
    return [_p(s.voxelVerts), _p(s.voxelVertsScan), _p(s.voxelOccupied), _p(s.voxelOccupiedScan), _p(s.compVoxelArray)]


def isosurface_lattice(one, fix_grid, vol, pos, norm, iso, dims, voxel, center, s, max_verts, vol_one, vol_two, isovalue1, isovalue2, iso1=0.0, iso2=0.0):
    act, tot = C.c_uint(0), C.c_uint(0)
    _ok(lib().ref_isosurface_lattice(int(one), int(fix_grid), _p(vol), _p(pos), _p(norm), F(iso), C.c_uint(dims[0]), C.c_uint(dims[1]), C.c_uint(dims[2]),
                                     F(voxel[0]), F(voxel[1]), F(voxel[2]), F(center[0]), F(center[1]), F(center[2]), *_scr(s), C.c_uint(max_verts),
                                     _p(vol_one), _p(vol_two), F(isovalue1), F(isovalue2), F(iso1), F(iso2), C.byref(act), C.byref(tot)))
    return act.value, tot.value


def isosurface_csg(fix_grid, pos, norm, iso, dims, voxel, center, s, max_verts, fixed_f, dynamic_f, lattice_f, iso1=0.2, iso2=0.3, obj_union=True,
                   obj_diff=False, obj_intersect=False, fixed=False, dynamic=False, make_region=False, topo_f=None):
    # classifyVoxel gathers all three fields unconditionally (MarchingCubes_kernel.cu:889-914)
    if lattice_f is None:
        lattice_f = _zeros_like_grid(dims, fixed_f)
    if dynamic_f is None:
        dynamic_f = _zeros_like_grid(dims, fixed_f)
    act, tot = C.c_uint(0), C.c_uint(0)
    _ok(lib().ref_isosurface_csg(int(fix_grid), _p(pos), _p(norm), F(iso), C.c_uint(dims[0]), C.c_uint(dims[1]), C.c_uint(dims[2]), F(voxel[0]), F(voxel[1]),
                                 F(voxel[2]), F(center[0]), F(center[1]), F(center[2]), *_scr(s), C.c_uint(max_verts), _p(fixed_f), _p(dynamic_f),
                                 _p(topo_f), _p(lattice_f), F(iso1), F(iso2), int(obj_union), int(obj_diff), int(obj_intersect), int(fixed), int(dynamic),
                                 int(make_region), C.byref(act), C.byref(tot)))
    return act.value, tot.value


def isosurface_region(fix_grid, pos, norm, iso, dims, voxel, center, s, max_verts, vol_topo, fixed_f, dynamic_f, make_region=False, show_region=False,
                      show_domain=False, triangle_data=None):
    act, tot = C.c_uint(0), C.c_uint(0)
    _ok(lib().ref_isosurface_region(int(fix_grid), _p(pos), _p(norm), F(iso), C.c_uint(dims[0]), C.c_uint(dims[1]), C.c_uint(dims[2]), F(voxel[0]),
                                    F(voxel[1]), F(voxel[2]), F(center[0]), F(center[1]), F(center[2]), *_scr(s), C.c_uint(max_verts), _p(vol_topo),
                                    _p(fixed_f), _p(dynamic_f), int(make_region), int(show_region), int(show_domain), _p(triangle_data),
                                    C.byref(act), C.byref(tot)))
    return act.value, tot.value


def isosurface_topo(with_disp_variant, pos, norm, iso, dims, voxel, center, s, max_verts, vol_topo, vol_two, isovalue1, d_result, disp=False, disp_two=None,
                    vol_one=None, d_solid=None):
    act, tot = C.c_uint(0), C.c_uint(0)
    _ok(lib().ref_isosurface_topo(int(with_disp_variant), _p(pos), _p(norm), F(iso), C.c_uint(dims[0]), C.c_uint(dims[1]), C.c_uint(dims[2]), F(voxel[0]),
                                  F(voxel[1]), F(voxel[2]), F(center[0]), F(center[1]), F(center[2]), *_scr(s), C.c_uint(max_verts), _p(vol_topo),
                                  _p(vol_one), _p(vol_two), _p(d_solid), F(isovalue1), _p(d_result), int(disp), _p(disp_two), C.byref(act), C.byref(tot)))
    return act.value, tot.value


def write_obj(pos, total_verts, filename):
    _ok(lib().ref_write_obj(_p(pos), C.c_uint(total_verts), filename.encode()))


# ---- SVL phase solve
def period_data(d_period, dims, d, mean, axis="z"):
    _ok(lib().ref_period_data(_p(d_period), dims[0], dims[1], dims[2], F(d[0]), F(d[1]), F(d[2]), F(mean[0]), F(mean[1]), F(mean[2]), ord(axis)))


def angle_data(d_theta, dims, d, mean, axis="z"):
    _ok(lib().ref_angle_data(_p(d_theta), dims[0], dims[1], dims[2], F(d[0]), F(d[1]), F(d[2]), F(mean[0]), F(mean[1]), F(mean[2]), ord(axis)))


def normalise_three(d_in, d_out, size, a1, b1):
    _ok(lib().ref_normalise_three(_p(d_in), _p(d_out), C.c_size_t(size), F(a1), F(b1)))


def finding_phi(d_phi, d_period, dims, ijk, d, latticetype="r", uniform_type=2, const_period=8.0, periods=(8.0, 8.0, 8.0), lcon=0.5, lcon_1=0.05, sinewave_zaxis=False):
    _ok(lib().ref_finding_phi(_p(d_phi), _p(d_period), dims[0], dims[1], dims[2], ijk[0], ijk[1], ijk[2], F(d[0]), F(d[1]), F(d[2]), ord(latticetype), uniform_type,
                              F(const_period), F(periods[0]), F(periods[1]), F(periods[2]), F(lcon), F(lcon_1), int(sinewave_zaxis)))


def cg(d_phi, dims, iters=500, end_res=0.01):
    fi, fr = C.c_int(0), C.c_float(0)
    _ok(lib().ref_cg(_p(d_phi), dims[0], dims[1], dims[2], iters, F(end_res), C.byref(fi), C.byref(fr)))
    return fi.value, fr.value
