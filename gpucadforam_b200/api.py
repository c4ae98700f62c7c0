"""Host-side mirror of the reference's C++ classes on top of the C ABI.

Class and method names follow the reference (`Isosurface::computeIsosurface_lattice`,
`Modelling::sphere_with_center`, `Gratings::GPU_buffer_normalise_four`, ... -- see
src/Isosurface.h, src/Modelling.h, src/lattice_files/Gratings.h, src/lattice_files/Fft_lattice.h,
src/File_output.h of the reference) so tests read like calls into the reference.  Device
buffers are torch CUDA tensors used as raw memory; every call goes straight to
libgpucad_b200.so.  No computation happens in Python.
"""
import ctypes as C

import torch

from . import _capi
from ._capi import Float3, PitchedPtr, Slab, Uint3

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _capi.load()
    return _lib


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, int):
        return C.c_void_p(t)
    assert t.is_cuda and t.is_contiguous(), "device buffers must be contiguous CUDA tensors"
    return C.c_void_p(t.data_ptr())


def _u3(v):
    return Uint3(int(v[0]), int(v[1]), int(v[2]))


def _f3(v):
    return Float3(float(v[0]), float(v[1]), float(v[2]))


class Context:
    """One gcb_ctx per device/stream (the reference uses process globals and the legacy stream)."""

    def __init__(self, device=0, stream=None, options=None):
        if not torch.cuda.is_available():
            raise RuntimeError("gpucadforam_b200 needs a CUDA device: the library has no CPU path")
        self.device = device
        self._h = C.c_void_p()
        torch.cuda.set_device(device)
        s = C.c_void_p(stream) if stream else None
        rc = lib().gcb_create(C.byref(self._h), device, s)
        if rc != 0:
            raise RuntimeError("gcb_create failed (%d)" % rc)
        if options is not None:
            self.set_options(options)

    def check(self, rc):
        if rc != 0:
            raise RuntimeError("libgpucad_b200: " + lib().gcb_last_error(self._h).decode())

    def set_options(self, flags):
        self.check(lib().gcb_set_options(self._h, flags))

    def set_stream(self, stream):
        self.check(lib().gcb_set_stream(self._h, C.c_void_p(stream) if stream else None))

    def launch_count(self):
        return int(lib().gcb_launch_count(self._h))

    def reset_launch_count(self):
        lib().gcb_reset_launch_count(self._h)

    def enable_kernel_timing(self, on=True):
        self.check(lib().gcb_enable_kernel_timing(self._h, 1 if on else 0))

    def last_extract_kernel_ms(self):
        return float(lib().gcb_last_extract_kernel_ms(self._h))

    def last_field_kernel_ms(self):
        return float(lib().gcb_last_field_kernel_ms(self._h))

    def close(self):
        if self._h:
            lib().gcb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Scratch:
    """The five caller-owned `uint[cells]` arrays of initMC_two (main.cu:2141-2151)."""

    def __init__(self, num_voxels, device="cuda"):
        mk = lambda: torch.zeros(max(int(num_voxels), 1), dtype=torch.int32, device=device)
        self.voxelVerts, self.voxelVertsScan, self.voxelOccupied, self.voxelOccupiedScan, self.compVoxelArray = mk(), mk(), mk(), mk(), mk()


class MeshBuffers:
    """`float4 pos[maxVerts]`, `float4 norm[maxVerts]` (Vulkan-exported in the app, main.cu:2736-2790)."""

    def __init__(self, max_verts, device="cuda"):
        self.max_verts = int(max_verts)
        self.pos = torch.zeros((self.max_verts, 4), dtype=torch.float32, device=device)
        self.norm = torch.zeros((self.max_verts, 4), dtype=torch.float32, device=device)


def grid_desc(nx, ny, nz):
    """gridSize, gridSizeShift, gridSizeMask, numVoxels as initMC_two builds them (main.cu:2125-2139)."""
    return (nx, ny, nz), (1, nx - 1, (nx - 1) * (ny - 1)), (nx - 1, ny - 1, nz - 1), (nx - 1) * (ny - 1) * (nz - 1)


class Isosurface:
    def __init__(self, ctx):
        self.ctx = ctx
        ctx.check(lib().gcb_allocateTextures_s(ctx._h, None, None))

    def computeIsosurface(self, pos, norm, isoValue, scratch, gridSize, voxelSize, gridcenter, maxVerts, primitive_fixed, primitive_dynamic,
                          lattice_field=None, iso1=0.2, iso2=0.3, obj_union=True, obj_diff=False, obj_intersect=False, fixed=False, dynamic=False,
                          make_region=False, topo_field=None):
        gs, sh, mk, nv = grid_desc(*gridSize)
        act, tot, nf = C.c_uint(0), C.c_uint(0), C.c_size_t(0)
        s = scratch
        self.ctx.check(lib().gcb_computeIsosurface(
            self.ctx._h, None, _u3(gs), _ptr(pos), _ptr(norm), isoValue, nv, _ptr(s.voxelVerts), _ptr(s.voxelVertsScan), _ptr(s.voxelOccupied),
            _ptr(s.voxelOccupiedScan), _u3(gs), _u3(sh), _u3(mk), _f3(voxelSize), _f3(gridcenter), C.byref(act), C.byref(tot), _ptr(s.compVoxelArray),
            maxVerts, _ptr(primitive_fixed), _ptr(primitive_dynamic), _ptr(topo_field), _ptr(lattice_field), iso1, iso2, int(obj_union), int(obj_diff),
            int(obj_intersect), 1, 0, 0, int(fixed), int(dynamic), int(make_region), C.byref(nf)))
        return act.value, tot.value, nf.value

    def computeIsosurface_region(self, pos, norm, isoValue, scratch, gridSize, voxelSize, gridcenter, maxVerts, vol_topo, primitive_fixed,
                                 primitive_dynamic, make_region=False, show_region=False, show_domain=False, triangle_data=None):
        """Isosurface::computeIsosurface_region (Isosurface.cu:150-239).  triangle_data: device buffer of 64-byte triangle_metadata
        records (torch tensor of shape [n, 16] int32), written when show_region is set."""
        gs, sh, mk, nv = grid_desc(*gridSize)
        act, tot = C.c_uint(0), C.c_uint(0)
        s = scratch
        self.ctx.check(lib().gcb_computeIsosurface_region(
            self.ctx._h, _ptr(pos), _ptr(norm), isoValue, nv, _ptr(s.voxelVerts), _ptr(s.voxelVertsScan), _ptr(s.voxelOccupied), _ptr(s.voxelOccupiedScan),
            _u3(gs), _u3(sh), _u3(mk), _f3(voxelSize), _f3(gridcenter), C.byref(act), C.byref(tot), _ptr(s.compVoxelArray), maxVerts, _ptr(vol_topo),
            _ptr(primitive_fixed), _ptr(primitive_dynamic), None, None, 0.0, 0.0, 1, 0, 0, 1, 0, 0, 0, 0, int(make_region), int(show_region),
            int(show_domain), _ptr(triangle_data)))
        return act.value, tot.value

    def computeIsosurface_lattice(self, vol, pos, norm, isoValue, scratch, gridSize, voxelSize, gridcenter, maxVerts, vol_one, vol_two, isovalue1,
                                  isovalue2, iso1=0.0, iso2=0.0):
        gs, sh, mk, nv = grid_desc(*gridSize)
        act, tot = C.c_uint(0), C.c_uint(0)
        s = scratch
        self.ctx.check(lib().gcb_computeIsosurface_lattice(
            self.ctx._h, _ptr(vol), _ptr(pos), _ptr(norm), isoValue, nv, _ptr(s.voxelVerts), _ptr(s.voxelVertsScan), _ptr(s.voxelOccupied),
            _ptr(s.voxelOccupiedScan), _u3(gs), _u3(sh), _u3(mk), _f3(voxelSize), _f3(gridcenter), C.byref(act), C.byref(tot), _ptr(s.compVoxelArray),
            maxVerts, _ptr(vol_one), _ptr(vol_two), isovalue1, isovalue2, iso1, iso2))
        return act.value, tot.value

    def computeIsosurface_latticeone(self, vol, pos, norm, isoValue, scratch, gridSize, voxelSize, gridcenter, maxVerts, vol_one, isovalue1, isovalue2):
        gs, sh, mk, nv = grid_desc(*gridSize)
        act, tot = C.c_uint(0), C.c_uint(0)
        s = scratch
        self.ctx.check(lib().gcb_computeIsosurface_latticeone(
            self.ctx._h, _ptr(vol), _ptr(pos), _ptr(norm), isoValue, nv, _ptr(s.voxelVerts), _ptr(s.voxelVertsScan), _ptr(s.voxelOccupied),
            _ptr(s.voxelOccupiedScan), _u3(gs), _u3(sh), _u3(mk), _f3(voxelSize), _f3(gridcenter), C.byref(act), C.byref(tot), _ptr(s.compVoxelArray),
            maxVerts, _ptr(vol_one), isovalue1, isovalue2))
        return act.value, tot.value

    def computeIsosurface_2(self, pos, norm, isoValue, scratch, gridSize, voxelSize, gridcenter, maxVerts, vol_topo, vol_two, isovalue1, d_result,
                            vol_one=None, d_solid=None):
        gs, sh, mk, nv = grid_desc(*gridSize)
        act, tot = C.c_uint(0), C.c_uint(0)
        s = scratch
        self.ctx.check(lib().gcb_computeIsosurface_2(
            self.ctx._h, _ptr(pos), _ptr(norm), isoValue, nv, _ptr(s.voxelVerts), _ptr(s.voxelVertsScan), _ptr(s.voxelOccupied), _ptr(s.voxelOccupiedScan),
            _u3(gs), _u3(sh), _u3(mk), _f3(voxelSize), _f3(gridcenter), C.byref(act), C.byref(tot), _ptr(s.compVoxelArray), maxVerts, _ptr(vol_topo),
            _ptr(vol_one), _ptr(vol_two), _ptr(d_solid), isovalue1, _ptr(d_result), None))
        return act.value, tot.value

    def computeIsosurface_topo(self, pos, norm, isoValue, scratch, gridSize, voxelSize, gridcenter, maxVerts, vol_topo, vol_two, isovalue1, d_result,
                               disp=False, disp_two=None, vol_one=None, d_solid=None):
        gs, sh, mk, nv = grid_desc(*gridSize)
        act, tot = C.c_uint(0), C.c_uint(0)
        s = scratch
        self.ctx.check(lib().gcb_computeIsosurface_topo(
            self.ctx._h, _ptr(pos), _ptr(norm), isoValue, nv, _ptr(s.voxelVerts), _ptr(s.voxelVertsScan), _ptr(s.voxelOccupied), _ptr(s.voxelOccupiedScan),
            _u3(gs), _u3(sh), _u3(mk), _f3(voxelSize), _f3(gridcenter), C.byref(act), C.byref(tot), _ptr(s.compVoxelArray), maxVerts, _ptr(vol_topo),
            _ptr(vol_one), _ptr(vol_two), _ptr(d_solid), isovalue1, _ptr(d_result), None, int(disp), _ptr(disp_two)))
        return act.value, tot.value

    def copy_parameter(self, isoValue, gridSize, voxelSize, vol_one, vol_two, vol_lattice=None, fixed=False, dynamic=False, iso1=0.2, iso2=0.3,
                       obj_union=True, obj_diff=False, obj_intersect=False):
        gs, sh, mk, nv = grid_desc(*gridSize)
        self.ctx.check(lib().gcb_copy_parameter(self.ctx._h, None, isoValue, _u3(gs), _u3(sh), _u3(mk), _f3(voxelSize), nv, _ptr(vol_one), _ptr(vol_two),
                                                _ptr(vol_lattice), int(fixed), int(dynamic), iso1, iso2, int(obj_union), int(obj_diff), int(obj_intersect)))

    def patch_topo_field(self, d_vec1, Nx, Ny, Nz, vol_one):
        self.ctx.check(lib().gcb_patch_topo_field(self.ctx._h, _ptr(d_vec1), Nx, Ny, Nz, _ptr(vol_one)))


class Modelling:
    def __init__(self, ctx):
        self.ctx = ctx

    def distance_from_line(self, data_1, center, axis, radius_1, thickness_radial, thickness_axial, Nx, Ny, Nz, dx, dy, dz, cylind_disc_selected):
        self.ctx.check(lib().gcb_distance_from_line(self.ctx._h, _ptr(data_1), _f3(center), _f3(axis), radius_1, thickness_radial, thickness_axial, Nx, Ny,
                                                    Nz, dx, dy, dz, int(cylind_disc_selected)))

    def sphere_with_center(self, data_1, center, radius_1, thickness_wall, Nx, Ny, Nz, dx, dy, dz, sphere_shell_selected):
        self.ctx.check(lib().gcb_sphere_with_center(self.ctx._h, _ptr(data_1), _f3(center), radius_1, thickness_wall, Nx, Ny, Nz, dx, dy, dz,
                                                    int(sphere_shell_selected)))

    def cuboid(self, data_1, center, angles, x_width, y_width, z_width, Nx, Ny, Nz, dx, dy, dz):
        self.ctx.check(lib().gcb_cuboid(self.ctx._h, _ptr(data_1), _f3(center), _f3(angles), x_width, y_width, z_width, Nx, Ny, Nz, dx, dy, dz))

    def cuboid_shell(self, data_1, center, angles, x_width, y_width, z_width, thickness, Nx, Ny, Nz, dx, dy, dz):
        self.ctx.check(lib().gcb_cuboid_shell(self.ctx._h, _ptr(data_1), _f3(center), _f3(angles), x_width, y_width, z_width, thickness, Nx, Ny, Nz, dx, dy,
                                              dz))

    def torus_with_center(self, data_1, center, angles, torus_radius, torus_circle_radius, Nx, Ny, Nz, dx, dy, dz):
        self.ctx.check(lib().gcb_torus_with_center(self.ctx._h, _ptr(data_1), _f3(center), _f3(angles), torus_radius, torus_circle_radius, Nx, Ny, Nz, dx, dy,
                                                   dz))

    def cone_with_base_radius_height(self, data_1, center, angles, base_radius, cone_height, Nx, Ny, Nz, dx, dy, dz):
        self.ctx.check(lib().gcb_cone_with_base_radius_height(self.ctx._h, _ptr(data_1), _f3(center), _f3(angles), base_radius, cone_height, Nx, Ny, Nz, dx,
                                                              dy, dz))

    def cone_frustum(self, data_1, center, angles, top_radius, bottom_radius, cone_frustum_height, Nx, Ny, Nz, dx, dy, dz):
        self.ctx.check(lib().gcb_cone_frustum(self.ctx._h, _ptr(data_1), _f3(center), _f3(angles), top_radius, bottom_radius, cone_frustum_height, Nx, Ny, Nz,
                                              dx, dy, dz))

    def pyramid_frustum(self, data_1, center, angles, x_width_base, x_width_top, y_height, z_width_base, z_width_top, Nx, Ny, Nz, dx, dy, dz):
        self.ctx.check(lib().gcb_pyramid_frustum(self.ctx._h, _ptr(data_1), _f3(center), _f3(angles), x_width_base, x_width_top, y_height, z_width_base,
                                                 z_width_top, Nx, Ny, Nz, dx, dy, dz))


class Fft_lattice:
    def __init__(self, ctx):
        self.ctx = ctx

    def create_lattice(self, d_latticevol, NX, NY, NZ, size, lattice_type_index):
        self.ctx.check(lib().gcb_create_lattice(self.ctx._h, _ptr(d_latticevol), NX, NY, NZ, size, lattice_type_index))


class Gratings:
    """Gratings : Interpolations (texture = context-owned control grid)."""

    def __init__(self, ctx):
        self.ctx = ctx

    def GPU_buffer_normalise_buffer(self, d_vec1, d_vec2, n):
        self.ctx.check(lib().gcb_GPU_buffer_normalise_buffer(self.ctx._h, _ptr(d_vec1), _ptr(d_vec2), n))

    def GPU_buffer_normalise_four(self, dataone, datatwo, datathree, size, Nx, Ny, Nz, isoval_1, isoval_2):
        self.ctx.check(lib().gcb_GPU_buffer_normalise_four(self.ctx._h, _ptr(dataone), _ptr(datatwo), _ptr(datathree), size, Nx, Ny, Nz, isoval_1, isoval_2))

    def GPU_buffer_normalise_three(self, dataone, datatwo, size, a1, b1):
        self.ctx.check(lib().gcb_GPU_buffer_normalise_three(self.ctx._h, _ptr(dataone), _ptr(datatwo), size, a1, b1))

    def period_data(self, d_period, NX, NY, NZ, dx, dy, dz, mean_x, mean_y, mean_z, axis="z"):
        self.ctx.check(lib().gcb_period_data(self.ctx._h, _ptr(d_period), NX, NY, NZ, dx, dy, dz, mean_x, mean_y, mean_z, axis.encode()))

    def angle_data(self, d_theta, NX, NY, NZ, dx, dy, dz, mean_x, mean_y, mean_z, axis="z"):
        self.ctx.check(lib().gcb_angle_data(self.ctx._h, _ptr(d_theta), NX, NY, NZ, dx, dy, dz, mean_x, mean_y, mean_z, axis.encode()))

    def grating(self, dvol, NX2, NY2, NZ2, dx2, dy2, dz2):
        self.ctx.check(lib().gcb_grating(self.ctx._h, _ptr(dvol), NX2, NY2, NZ2, dx2, dy2, dz2))

    def refine(self, dvol, NX2, NY2, NZ2, dx, dy, dz):
        self.ctx.check(lib().gcb_refine(self.ctx._h, _ptr(dvol), NX2, NY2, NZ2, dx, dy, dz))

    def svl(self, d_svl, d_grating, NX, NY, NZ, indxx, data_fft):
        self.ctx.check(lib().gcb_svl(self.ctx._h, _ptr(d_svl), _ptr(d_grating), NX, NY, NZ, indxx, _ptr(data_fft)))

    def topo_field(self, topo_field, isosurf, volfrac, NX, NY, NZ):
        self.ctx.check(lib().gcb_topo_field(self.ctx._h, _ptr(topo_field), _ptr(isosurf), volfrac, NX, NY, NZ))

    def primitive_field(self, primitive_field, primitive_active, isosurf, isoval, fixed, active, NX, NY, NZ):
        self.ctx.check(lib().gcb_primitive_field(self.ctx._h, _ptr(primitive_field), _ptr(primitive_active), _ptr(isosurf), isoval, int(fixed), int(active),
                                                 NX, NY, NZ))

    # Interpolations
    def setupTexture(self, dx, dy, dz):
        self.ctx.check(lib().gcb_setupTexture(self.ctx._h, dx, dy, dz))

    def copytotexture(self, d_phi, pitched, NX, NY, NZ):
        self.ctx.check(lib().gcb_copytotexture(self.ctx._h, _ptr(d_phi), pitched, NX, NY, NZ))

    def updateTexture(self, pitched):
        self.ctx.check(lib().gcb_updateTexture(self.ctx._h, pitched))

    def deleteTexture(self):
        self.ctx.check(lib().gcb_deleteTexture(self.ctx._h))

    @staticmethod
    def pitched(buf, nx, ny):
        """cudaPitchedPtr over a dense torch buffer (pitch = nx*4)."""
        return PitchedPtr(buf.data_ptr(), nx * 4, nx * 4, ny)


def finding_phi(ctx, d_phi, d_period, dims, ijk, d, latticetype="r", uniform_type=2, const_period=8.0, periods=(8.0, 8.0, 8.0), lcon=0.5, lcon_1=0.05,
                sinewave_zaxis=False):
    """Gratings::finding_phi for one harmonic (i, j, k) (Gratings.cu:1015-1025)."""
    ctx.check(lib().gcb_finding_phi(ctx._h, _ptr(d_phi), _ptr(d_period), dims[0], dims[1], dims[2], ijk[0], ijk[1], ijk[2], d[0], d[1], d[2],
                                    latticetype.encode(), uniform_type, const_period, periods[0], periods[1], periods[2], lcon, lcon_1, int(sinewave_zaxis)))


def GPUCG_lattice(ctx, d_phi, dims, iters=500, end_res=0.01):
    """Gratings::GPUCG_lattice (Gratings.cu:875-974): d_phi holds the right-hand side on entry, the solution on return."""
    fi, fr = C.c_int(0), C.c_float(0)
    ctx.check(lib().gcb_GPUCG_lattice(ctx._h, _ptr(d_phi), dims[0], dims[1], dims[2], iters, 1, end_res, C.byref(fi), C.byref(fr)))
    return fi.value, fr.value


def svl_phase_solve(ctx, d_phi_all, d_period, harmonics, dims, d, latticetype="r", uniform_type=2, const_period=8.0, periods=(8.0, 8.0, 8.0), lcon=0.5,
                    lcon_1=0.05, sinewave_zaxis=False, iters=500, end_res=0.01):
    """All harmonics at once: right-hand sides + batched CG.  Returns (FinalIter list, FinalRes list)."""
    nh = len(harmonics)
    flat = (C.c_int * (3 * nh))(*[int(v) for h in harmonics for v in h])
    fi, fr = (C.c_int * nh)(), (C.c_float * nh)()
    ctx.check(lib().gcb_svl_phase_solve(ctx._h, _ptr(d_phi_all), _ptr(d_period), nh, flat, dims[0], dims[1], dims[2], d[0], d[1], d[2], latticetype.encode(),
                                        uniform_type, const_period, periods[0], periods[1], periods[2], lcon, lcon_1, int(sinewave_zaxis), iters, end_res, fi, fr))
    return list(fi), list(fr)


def unit_lattice_spectrum(ctx, d_unit_cell, Nxu, Nyu, Nzu, range_st=2):
    """lattice_data of Multitopo::unit_lattice (main.cu:3577-3706): complex64 tensor [(2*range_st+1)^3] on the device."""
    side = 2 * range_st + 1
    out = torch.zeros(side ** 3, 2, dtype=torch.float32, device=d_unit_cell.device)
    ctx.check(lib().gcb_unit_lattice_spectrum(ctx._h, _ptr(d_unit_cell), Nxu, Nyu, Nzu, range_st, _ptr(out)))
    return torch.view_as_complex(out)


class File_output:
    def __init__(self, ctx):
        self.ctx = ctx

    def file_write_obj(self, d_pos, totalVerts, filename):
        self.ctx.check(lib().gcb_file_write_obj(self.ctx._h, _ptr(d_pos), totalVerts, filename.encode()))


# ---------------------------------------------------------------- fused entry points
def _coef_array(coef):
    flat = [float(v) for pair in coef for v in pair]
    return (C.c_float * len(flat))(*flat)


def minmax(ctx, d_in, n=None):
    """(min(0, min f), max(0, max f)) on the host, the reference reduction's semantics (Gratings.cu:1394-1495)."""
    lo, hi = C.c_float(0), C.c_float(0)
    ctx.check(lib().gcb_minmax(ctx._h, _ptr(d_in), int(d_in.numel() if n is None else n), C.byref(lo), C.byref(hi)))
    return lo.value, hi.value


def svl_field(ctx, d_svl, d_phi, coef, cdims, fdims, d, slab=(0, 0), cz0=0, accumulate=False, d_minmax=None):
    cx, cy, czl = cdims
    nx2, ny2, nz2l = fdims
    ctx.check(lib().gcb_svl_field(ctx._h, _ptr(d_svl), _ptr(d_phi), len(coef), _coef_array(coef), cx, cy, czl, cz0, nx2, ny2, nz2l,
                                  Slab(slab[0], slab[1] or nz2l), d[0], d[1], d[2], int(accumulate), _ptr(d_minmax)))


def svl_field_host(ctx, d_svl, h_phi, d_phi_scratch, coef, cdims, fdims, d, slab=(0, 0), cz0=0, d_minmax=None):
    """h_phi: pinned (or pageable) HOST tensor [nh, czl, cy, cx]; batched upload overlapped with the field kernel."""
    assert not h_phi.is_cuda and h_phi.is_contiguous()
    cx, cy, czl = cdims
    nx2, ny2, nz2l = fdims
    ctx.check(lib().gcb_svl_field_host(ctx._h, _ptr(d_svl), C.c_void_p(h_phi.data_ptr()), _ptr(d_phi_scratch), len(coef), _coef_array(coef), cx, cy, czl, cz0,
                                       nx2, ny2, nz2l, Slab(slab[0], slab[1] or nz2l), d[0], d[1], d[2], _ptr(d_minmax)))


def extract_band_raw(ctx, d_field, a, b, isoValue, isovalue1, isovalue2, gridSizeLocal, voxelSize, gridcenter, pos, norm, maxVerts, slab=(0, 0),
                     comp=None, count_only=False):
    act, tot = C.c_ulonglong(0), C.c_ulonglong(0)
    ctx.check(lib().gcb_extract_band_raw(ctx._h, _ptr(d_field), a, b, isoValue, isovalue1, isovalue2, _u3(gridSizeLocal),
                                         Slab(slab[0], slab[1] or gridSizeLocal[2]), _f3(voxelSize), _f3(gridcenter), _ptr(pos), _ptr(norm), maxVerts,
                                         _ptr(comp), int(count_only), C.byref(act), C.byref(tot)))
    return act.value, tot.value


def extract_band_raw_dev(ctx, d_field, d_minmax, isoValue, isovalue1, isovalue2, gridSizeLocal, voxelSize, gridcenter, pos, norm, maxVerts, slab=(0, 0),
                         comp=None, count_only=False):
    """extract_band_raw with the normalisation range {min, max} read from device memory (2-float CUDA tensor)."""
    act, tot = C.c_ulonglong(0), C.c_ulonglong(0)
    ctx.check(lib().gcb_extract_band_raw_dev(ctx._h, _ptr(d_field), _ptr(d_minmax), isoValue, isovalue1, isovalue2, _u3(gridSizeLocal),
                                             Slab(slab[0], slab[1] or gridSizeLocal[2]), _f3(voxelSize), _f3(gridcenter), _ptr(pos), _ptr(norm), maxVerts,
                                             _ptr(comp), int(count_only), C.byref(act), C.byref(tot)))
    return act.value, tot.value


def band_lattice_from_raw(ctx, d_raw_field, gridSize, isoValue, isovalue1, isovalue2, voxelSize, gridcenter, pos, norm, maxVerts, comp=None):
    """normalise_buffer -> normalise_four -> latticeone on a raw field, fused (gcb_band_lattice_from_raw): (active, verts, (a, b, a2, b2))."""
    act, tot = C.c_ulonglong(0), C.c_ulonglong(0)
    rg = (C.c_float * 4)()
    ctx.check(lib().gcb_band_lattice_from_raw(ctx._h, _ptr(d_raw_field), _u3(gridSize), isoValue, isovalue1, isovalue2, _f3(voxelSize), _f3(gridcenter), _ptr(pos),
                                              _ptr(norm), maxVerts, _ptr(comp), C.byref(act), C.byref(tot), rg))
    return act.value, tot.value, tuple(rg)


def tpms_lattice(ctx, d_field_scratch, lattice_type_index, gridSize, isoValue, isovalue1, isovalue2, voxelSize, gridcenter, pos, norm, maxVerts, comp=None):
    """BASELINE config 1 in one call (gcb_tpms_lattice): create_lattice + both normalisations + latticeone."""
    act, tot = C.c_ulonglong(0), C.c_ulonglong(0)
    rg = (C.c_float * 4)()
    ctx.check(lib().gcb_tpms_lattice(ctx._h, _ptr(d_field_scratch), int(lattice_type_index), _u3(gridSize), isoValue, isovalue1, isovalue2, _f3(voxelSize),
                                     _f3(gridcenter), _ptr(pos), _ptr(norm), maxVerts, _ptr(comp), C.byref(act), C.byref(tot), rg))
    return act.value, tot.value, tuple(rg)


def density_surface(ctx, d_coarse, cdims, d_density_fine, fdims, d, isoValue, voxelSize, gridcenter, pos, norm, maxVerts, comp=None):
    """BASELINE config 5 in one call (gcb_density_surface): 2x upsample of the coarse density + computeIsosurface_2 semantics."""
    act, tot = C.c_ulonglong(0), C.c_ulonglong(0)
    ctx.check(lib().gcb_density_surface(ctx._h, _ptr(d_coarse), cdims[0], cdims[1], cdims[2], _ptr(d_density_fine), fdims[0], fdims[1], fdims[2], d[0], d[1], d[2],
                                        isoValue, _f3(voxelSize), _f3(gridcenter), _ptr(pos), _ptr(norm), maxVerts, _ptr(comp), C.byref(act), C.byref(tot)))
    return act.value, tot.value


# kind -> (gcb_primitive_kind, names of the scalar parameters in the order of the primitive's own entry point, name of the aux vector, flag name)
_PRIM_KINDS = {
    "sphere": (0, ("radius", "thickness"), None, "shell"),
    "line": (1, ("radius", "thickness_radial", "thickness_axial"), "axis", "disc"),
    "cuboid": (2, ("xw", "yw", "zw"), "angles", None),
    "cuboid_shell": (3, ("xw", "yw", "zw", "thickness"), "angles", None),
    "torus": (4, ("torus_radius", "circle_radius"), "angles", None),
    "cone": (5, ("base_radius", "height"), "angles", None),
    "cone_frustum": (6, ("top_radius", "bottom_radius", "height"), "angles", None),
    "pyramid_frustum": (7, ("x_base", "x_top", "y_height", "z_base", "z_top"), "angles", None),
}


def csg_retain_primitive(ctx, kind, vol_one, d_field, dims, d, isoValue, obj_union=True, obj_diff=False, obj_intersect=False, center=(0.0, 0.0, 0.0), **kw):
    """Primitive + Isosurface::copy_parameter in one call (gcb_csg_retain_primitive).  d_field may be None where the primitive is evaluated
    inside the retain kernel (sphere, cuboid, cuboid_shell on rows that are a multiple of four points)."""
    code, names, aux_name, flag_name = _PRIM_KINDS[kind]
    params = (C.c_float * len(names))(*[float(kw[n]) for n in names])
    aux = kw.get(aux_name, (0.0, 0.0, 0.0)) if aux_name else (0.0, 0.0, 0.0)
    flag = int(bool(kw.get(flag_name, False))) if flag_name else 0
    ctx.check(lib().gcb_csg_retain_primitive(ctx._h, code, _f3(center), _f3(aux), params, len(names), flag, _ptr(d_field), _ptr(vol_one), dims[0], dims[1],
                                             dims[2], d[0], d[1], d[2], isoValue, int(obj_union), int(obj_diff), int(obj_intersect)))


def svl_lattice_host_submit(ctx, slot, h_phi, d_phi_scratch, d_svl_scratch, coef, cdims, fdims, d, isoValue, isovalue1, isovalue2, voxelSize, gridcenter, pos,
                            norm, maxVerts):
    """Enqueue one host-input job in pipeline slot 0 / 1 (gcb_svl_lattice_host_submit); returns immediately."""
    assert not h_phi.is_cuda and h_phi.is_contiguous()
    ctx.check(lib().gcb_svl_lattice_host_submit(ctx._h, int(slot), C.c_void_p(h_phi.data_ptr()), _ptr(d_phi_scratch), _ptr(d_svl_scratch), len(coef),
                                                _coef_array(coef), cdims[0], cdims[1], cdims[2], fdims[0], fdims[1], fdims[2], d[0], d[1], d[2], isoValue,
                                                isovalue1, isovalue2, _f3(voxelSize), _f3(gridcenter), _ptr(pos), _ptr(norm), maxVerts))


def svl_slab_host_submit_field(ctx, slot, h_phi, d_phi_scratch, d_svl_scratch, coef, cdims, fdims, d, slab, cz0, d_minmax):
    """First half of a sharded pipeline job (gcb_svl_slab_host_submit_field): upload + field of this rank's slab, local {min, max} -> d_minmax."""
    assert not h_phi.is_cuda and h_phi.is_contiguous()
    ctx.check(lib().gcb_svl_slab_host_submit_field(ctx._h, int(slot), C.c_void_p(h_phi.data_ptr()), _ptr(d_phi_scratch), _ptr(d_svl_scratch), len(coef),
                                                   _coef_array(coef), cdims[0], cdims[1], cdims[2], cz0, fdims[0], fdims[1], fdims[2], Slab(slab[0], slab[1] or fdims[2]),
                                                   d[0], d[1], d[2], _ptr(d_minmax)))


def svl_slab_host_submit_extract(ctx, slot, d_svl_scratch, d_ab, isoValue, isovalue1, isovalue2, gridSizeLocal, voxelSize, gridcenter, pos, norm, maxVerts,
                                 slab=(0, 0)):
    """Second half (gcb_svl_slab_host_submit_extract): extraction with the range over all ranks read from d_ab; svl_lattice_host_wait completes it."""
    ctx.check(lib().gcb_svl_slab_host_submit_extract(ctx._h, int(slot), _ptr(d_svl_scratch), _ptr(d_ab), isoValue, isovalue1, isovalue2, _u3(gridSizeLocal),
                                                     Slab(slab[0], slab[1] or gridSizeLocal[2]), _f3(voxelSize), _f3(gridcenter), _ptr(pos), _ptr(norm), maxVerts))


def svl_lattice_host_wait(ctx, slot):
    """Block until the job in `slot` is complete: (activeVoxels, totalVerts, (min, max))."""
    act, tot = C.c_ulonglong(0), C.c_ulonglong(0)
    mm = (C.c_float * 2)()
    ctx.check(lib().gcb_svl_lattice_host_wait(ctx._h, int(slot), C.byref(act), C.byref(tot), mm))
    return act.value, tot.value, (mm[0], mm[1])


def svl_lattice(ctx, d_svl_scratch, d_phi, coef, cdims, fdims, d, isoValue, isovalue1, isovalue2, voxelSize, gridcenter, pos, norm, maxVerts):
    act, tot = C.c_ulonglong(0), C.c_ulonglong(0)
    mm = (C.c_float * 2)()
    ctx.check(lib().gcb_svl_lattice(ctx._h, _ptr(d_svl_scratch), _ptr(d_phi), len(coef), _coef_array(coef), cdims[0], cdims[1], cdims[2], fdims[0], fdims[1],
                                    fdims[2], d[0], d[1], d[2], isoValue, isovalue1, isovalue2, _f3(voxelSize), _f3(gridcenter), _ptr(pos), _ptr(norm),
                                    maxVerts, C.byref(act), C.byref(tot), mm))
    return act.value, tot.value, (mm[0], mm[1])


def svl_lattice_host(ctx, h_phi, d_phi_scratch, d_svl_scratch, coef, cdims, fdims, d, isoValue, isovalue1, isovalue2, voxelSize, gridcenter, pos, norm,
                     maxVerts):
    """h_phi: pinned (or pageable) HOST torch tensor holding the control grids."""
    assert not h_phi.is_cuda and h_phi.is_contiguous()
    act, tot = C.c_ulonglong(0), C.c_ulonglong(0)
    mm = (C.c_float * 2)()
    ctx.check(lib().gcb_svl_lattice_host(ctx._h, C.c_void_p(h_phi.data_ptr()), _ptr(d_phi_scratch), _ptr(d_svl_scratch), len(coef), _coef_array(coef),
                                         cdims[0], cdims[1], cdims[2], fdims[0], fdims[1], fdims[2], d[0], d[1], d[2], isoValue, isovalue1, isovalue2,
                                         _f3(voxelSize), _f3(gridcenter), _ptr(pos), _ptr(norm), maxVerts, C.byref(act), C.byref(tot), mm))
    return act.value, tot.value, (mm[0], mm[1])
