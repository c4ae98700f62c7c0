// gpucad_host.hpp -- C++ host mirror of the reference's classes on top of the C ABI (include/gpucad_b200.h).
//
// A maintainer of GPUCADforAM switches the hot path to this engine by including this header instead of
// Isosurface.h / Modelling.h / lattice_files/{Fft_lattice,Gratings}.h / File_output.h and linking
// libgpucad_b200.so: class names, method names, parameter order and buffer layouts are the reference's
// (src/Isosurface.h:10-77, src/Modelling.h:10-50, src/lattice_files/Fft_lattice.h:11-24,
// src/lattice_files/Gratings.h:10-79, src/Interpolations.h:8-42, src/File_output.h:38), so the call sites in
// src/main.cu compile unchanged.  Error policy is the reference's: print and exit(EXIT_FAILURE)
// (commons/helper_cuda.h:583-612).  Header-only; needs <cuda_runtime_api.h> for uint3/float3/float4.
#pragma once
#include <cuda_runtime_api.h>
#include <cstdio>
#include <cstdlib>

#include "../../include/gpucad_b200.h"

#ifndef NTHREADS
#define NTHREADS 32
#endif
typedef unsigned int uint;
struct grid_points { int val = 0; float t_x = 0.0f, t_y = 0.0f, t_z = 0.0f; };  // src/MarchingCubes_kernel.h:12-18
struct triangle_metadata {                                                         // src/MarchingCubes_kernel.h:20-32
    uint index, voxel, l_index, edge_1, edge_2, edge_3, load_group;
    float3 centroid, normal, force_dir;
};
static_assert(sizeof(triangle_metadata) == sizeof(gcb_triangle_metadata), "triangle_metadata layout");

namespace gpucad {
inline gcb_ctx*& ctx_slot() { static gcb_ctx* c = nullptr; return c; }
// one process-wide context on the current device and the legacy default stream, like the reference's globals
inline gcb_ctx* ctx() {
    if (!ctx_slot()) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (gcb_create(&ctx_slot(), dev, nullptr) != 0) { fprintf(stderr, "gpucad_b200: no usable CUDA device\n"); exit(EXIT_FAILURE); }
    }
    return ctx_slot();
}
inline void check(int rc, const char* what) {
    if (rc != 0) { fprintf(stderr, "gpucad_b200: %s failed: %s\n", what, gcb_last_error(ctx())); exit(EXIT_FAILURE); }
}
inline gcb_uint3 u3(uint3 v) { return gcb_uint3{v.x, v.y, v.z}; }
inline gcb_float3 f3(float3 v) { return gcb_float3{v.x, v.y, v.z}; }
}  // namespace gpucad

class MarchingCubeCuda {
public:
    void allocateTextures_s(uint** d_triTable, uint** d_numVertsTable) { gpucad::check(gcb_allocateTextures_s(gpucad::ctx(), d_triTable, d_numVertsTable), "allocateTextures_s"); }
    void destroyAllTextureObjects() { gpucad::check(gcb_destroyAllTextureObjects(gpucad::ctx()), "destroyAllTextureObjects"); }
};

class Isosurface : public MarchingCubeCuda {
public:
    void copy_parameter(uint* voxel_verts, float isoValue, uint3 gridSize, uint3 gridSizeShift, uint3 gridSizeMask, float3 voxelSize, uint numVoxels,
                        grid_points* vol_one, float* vol_two, float* vol_lattice, bool fixed, bool dynamic, float iso1, float iso2, bool obj_union,
                        bool obj_diff, bool obj_intersect) {
        using namespace gpucad;
        check(gcb_copy_parameter(ctx(), voxel_verts, isoValue, u3(gridSize), u3(gridSizeShift), u3(gridSizeMask), f3(voxelSize), numVoxels,
                                 (gcb_grid_points*)vol_one, vol_two, vol_lattice, fixed, dynamic, iso1, iso2, obj_union, obj_diff, obj_intersect),
              "copy_parameter");
    }
    void computeIsosurface(float* vol, uint3 raster_grid, float4* pos, float4* norm, float isoValue, uint numVoxels, uint* d_voxelVerts,
                           uint* d_voxelVertsScan, uint* d_voxelOccupied, uint* d_voxelOccupiedScan, uint3 gridSize, uint3 gridSizeShift,
                           uint3 gridSizeMask, float3 voxelSize, float3 gridcenter, uint* activeVoxels, uint* totalVerts, uint* d_compVoxelArray,
                           uint maxVerts, grid_points* primitive_fixed, float* primitive_dynamic, float* topo_field, float* lattice_field, float iso1,
                           float iso2, bool obj_union, bool obj_diff, bool obj_intersect, bool primitive, bool topo, bool compute_lattice, bool fixed,
                           bool dynamic, bool make_region, size_t* nfacets) {
        using namespace gpucad;
        check(gcb_computeIsosurface(ctx(), vol, u3(raster_grid), pos, norm, isoValue, numVoxels, d_voxelVerts, d_voxelVertsScan, d_voxelOccupied,
                                    d_voxelOccupiedScan, u3(gridSize), u3(gridSizeShift), u3(gridSizeMask), f3(voxelSize), f3(gridcenter), activeVoxels,
                                    totalVerts, d_compVoxelArray, maxVerts, (gcb_grid_points*)primitive_fixed, primitive_dynamic, topo_field,
                                    lattice_field, iso1, iso2, obj_union, obj_diff, obj_intersect, primitive, topo, compute_lattice, fixed, dynamic,
                                    make_region, nfacets),
              "computeIsosurface");
    }
    void computeIsosurface_region(float4* pos, float4* norm, float isoValue, uint numVoxels, uint* d_voxelVerts, uint* d_voxelVertsScan, uint* d_voxelOccupied,
                                  uint* d_voxelOccupiedScan, uint3 gridSize, uint3 gridSizeShift, uint3 gridSizeMask, float3 voxelSize, float3 gridcenter,
                                  uint* activeVoxels, uint* totalVerts, uint* d_compVoxelArray, uint maxVerts, grid_points* vol_topo,
                                  grid_points* primitive_fixed, float* primitive_dynamic, float* topo_field, float* lattice_field, float iso1, float iso2,
                                  bool obj_union, bool obj_diff, bool obj_intersect, bool primitive, bool topo, bool compute_lattice, bool fixed, bool dynamic,
                                  bool make_region, bool show_region, bool show_domain, triangle_metadata* triangle_data) {
        using namespace gpucad;
        check(gcb_computeIsosurface_region(ctx(), pos, norm, isoValue, numVoxels, d_voxelVerts, d_voxelVertsScan, d_voxelOccupied, d_voxelOccupiedScan,
                                           u3(gridSize), u3(gridSizeShift), u3(gridSizeMask), f3(voxelSize), f3(gridcenter), activeVoxels, totalVerts,
                                           d_compVoxelArray, maxVerts, (gcb_grid_points*)vol_topo, (gcb_grid_points*)primitive_fixed, primitive_dynamic,
                                           topo_field, lattice_field, iso1, iso2, obj_union, obj_diff, obj_intersect, primitive, topo, compute_lattice, fixed,
                                           dynamic, make_region, show_region, show_domain, (gcb_triangle_metadata*)triangle_data),
              "computeIsosurface_region");
    }
    void computeIsosurface_2(float4* pos, float4* norm, float isoValue, uint numVoxels, uint* d_voxelVerts, uint* d_voxelVertsScan, uint* d_voxelOccupied,
                             uint* d_voxelOccupiedScan, uint3 gridSize, uint3 gridSizeShift, uint3 gridSizeMask, float3 voxelSize, float3 gridcenter,
                             uint* activeVoxels, uint* totalVerts, uint* d_compVoxelArray, uint maxVerts, grid_points* vol_topo, grid_points* vol_one,
                             float* vol_two, float* d_solid, float isovalue1, float* d_result, triangle_metadata* triangle_data) {
        using namespace gpucad;
        check(gcb_computeIsosurface_2(ctx(), pos, norm, isoValue, numVoxels, d_voxelVerts, d_voxelVertsScan, d_voxelOccupied, d_voxelOccupiedScan,
                                      u3(gridSize), u3(gridSizeShift), u3(gridSizeMask), f3(voxelSize), f3(gridcenter), activeVoxels, totalVerts,
                                      d_compVoxelArray, maxVerts, (gcb_grid_points*)vol_topo, (gcb_grid_points*)vol_one, vol_two, d_solid, isovalue1,
                                      d_result, (void*)triangle_data),
              "computeIsosurface_2");
    }
    void computeIsosurface_topo(float4* pos, float4* norm, float isoValue, uint numVoxels, uint* d_voxelVerts, uint* d_voxelVertsScan,
                                uint* d_voxelOccupied, uint* d_voxelOccupiedScan, uint3 gridSize, uint3 gridSizeShift, uint3 gridSizeMask,
                                float3 voxelSize, float3 gridcenter, uint* activeVoxels, uint* totalVerts, uint* d_compVoxelArray, uint maxVerts,
                                grid_points* vol_topo, grid_points* vol_one, float* vol_two, float* d_solid, float isovalue1, float* d_result,
                                triangle_metadata* triangle_data, bool disp, float4* disp_two) {
        using namespace gpucad;
        check(gcb_computeIsosurface_topo(ctx(), pos, norm, isoValue, numVoxels, d_voxelVerts, d_voxelVertsScan, d_voxelOccupied, d_voxelOccupiedScan,
                                         u3(gridSize), u3(gridSizeShift), u3(gridSizeMask), f3(voxelSize), f3(gridcenter), activeVoxels, totalVerts,
                                         d_compVoxelArray, maxVerts, (gcb_grid_points*)vol_topo, (gcb_grid_points*)vol_one, vol_two, d_solid, isovalue1,
                                         d_result, triangle_data, disp, disp_two),
              "computeIsosurface_topo");
    }
    void computeIsosurface_lattice(float* vol, float4* pos, float4* norm, float& isoValue, uint numVoxels, uint* d_voxelVerts, uint* d_voxelVertsScan,
                                   uint* d_voxelOccupied, uint* d_voxelOccupiedScan, uint3 gridSize, uint3 gridSizeShift, uint3 gridSizeMask,
                                   float3 voxelSize, float3 gridcenter, uint* activeVoxels, uint* totalVerts, uint* d_compVoxelArray, uint maxVerts,
                                   float* vol_one, float* vol_two, float isovalue1, float isovalue2, float iso1, float iso2) {
        using namespace gpucad;
        check(gcb_computeIsosurface_lattice(ctx(), vol, pos, norm, isoValue, numVoxels, d_voxelVerts, d_voxelVertsScan, d_voxelOccupied,
                                            d_voxelOccupiedScan, u3(gridSize), u3(gridSizeShift), u3(gridSizeMask), f3(voxelSize), f3(gridcenter),
                                            activeVoxels, totalVerts, d_compVoxelArray, maxVerts, vol_one, vol_two, isovalue1, isovalue2, iso1, iso2),
              "computeIsosurface_lattice");
    }
    void computeIsosurface_latticeone(float* vol, float4* pos, float4* norm, float& isoValue, uint numVoxels, uint* d_voxelVerts, uint* d_voxelVertsScan,
                                      uint* d_voxelOccupied, uint* d_voxelOccupiedScan, uint3 gridSize, uint3 gridSizeShift, uint3 gridSizeMask,
                                      float3 voxelSize, float3 gridcenter, uint* activeVoxels, uint* totalVerts, uint* d_compVoxelArray, uint maxVerts,
                                      float* vol_one, float isovalue1, float isovalue2) {
        using namespace gpucad;
        check(gcb_computeIsosurface_latticeone(ctx(), vol, pos, norm, isoValue, numVoxels, d_voxelVerts, d_voxelVertsScan, d_voxelOccupied,
                                               d_voxelOccupiedScan, u3(gridSize), u3(gridSizeShift), u3(gridSizeMask), f3(voxelSize), f3(gridcenter),
                                               activeVoxels, totalVerts, d_compVoxelArray, maxVerts, vol_one, isovalue1, isovalue2),
              "computeIsosurface_latticeone");
    }
    void patch_topo_field(float* d_vec1, int Nx, int Ny, int Nz, grid_points* vol_one) {
        gpucad::check(gcb_patch_topo_field(gpucad::ctx(), d_vec1, Nx, Ny, Nz, (gcb_grid_points*)vol_one), "patch_topo_field");
    }
};

class Modelling {
public:
    Modelling(int, int, int) {}
    void distance_from_line(float* data_1, float3 center, float3 axis, float radius_1, float thickness_radial, float thickness_axial, int Nx, int Ny, int Nz,
                            float dx, float dy, float dz, bool onetime) {
        using namespace gpucad;
        check(gcb_distance_from_line(ctx(), data_1, f3(center), f3(axis), radius_1, thickness_radial, thickness_axial, Nx, Ny, Nz, dx, dy, dz, onetime), "distance_from_line");
    }
    void sphere_with_center(float* data_1, float3 center, float radius_1, float thickness_wall, int Nx, int Ny, int Nz, float dx, float dy, float dz, bool onetime) {
        using namespace gpucad;
        check(gcb_sphere_with_center(ctx(), data_1, f3(center), radius_1, thickness_wall, Nx, Ny, Nz, dx, dy, dz, onetime), "sphere_with_center");
    }
    void cuboid(float* data_1, float3 center, float3 angles, float x_width, float y_width, float z_width, int Nx, int Ny, int Nz, float dx, float dy, float dz) {
        using namespace gpucad;
        check(gcb_cuboid(ctx(), data_1, f3(center), f3(angles), x_width, y_width, z_width, Nx, Ny, Nz, dx, dy, dz), "cuboid");
    }
    void cuboid_shell(float* data_1, float3 center, float3 angles, float x_width, float y_width, float z_width, float thickness, int Nx, int Ny, int Nz,
                      float dx, float dy, float dz) {
        using namespace gpucad;
        check(gcb_cuboid_shell(ctx(), data_1, f3(center), f3(angles), x_width, y_width, z_width, thickness, Nx, Ny, Nz, dx, dy, dz), "cuboid_shell");
    }
    void torus_with_center(float* data_1, float3 center, float3 angles, float torus_radius, float torus_circle_radius, int Nx, int Ny, int Nz, float dx,
                           float dy, float dz) {
        using namespace gpucad;
        check(gcb_torus_with_center(ctx(), data_1, f3(center), f3(angles), torus_radius, torus_circle_radius, Nx, Ny, Nz, dx, dy, dz), "torus_with_center");
    }
    void cone_with_base_radius_height(float* data_1, float3 center, float3 angles, float base_radius, float cone_height, int Nx, int Ny, int Nz, float dx,
                                      float dy, float dz) {
        using namespace gpucad;
        check(gcb_cone_with_base_radius_height(ctx(), data_1, f3(center), f3(angles), base_radius, cone_height, Nx, Ny, Nz, dx, dy, dz), "cone");
    }
    void cone_frustum(float* data_1, float3 center, float3 angles, float top_radius, float bottom_radius, float cone_frustum_height, int Nx, int Ny, int Nz,
                      float dx, float dy, float dz) {
        using namespace gpucad;
        check(gcb_cone_frustum(ctx(), data_1, f3(center), f3(angles), top_radius, bottom_radius, cone_frustum_height, Nx, Ny, Nz, dx, dy, dz), "cone_frustum");
    }
    void pyramid_frustum(float* data_1, float3 center, float3 angles, float x_width_base, float x_width_top, float y_height, float z_width_base,
                         float z_width_top, int Nx, int Ny, int Nz, float dx, float dy, float dz) {
        using namespace gpucad;
        check(gcb_pyramid_frustum(ctx(), data_1, f3(center), f3(angles), x_width_base, x_width_top, y_height, z_width_base, z_width_top, Nx, Ny, Nz, dx, dy, dz),
              "pyramid_frustum");
    }
};

class Fft_lattice {
public:
    void create_lattice(float* d_latticevol, uint NX, uint NY, uint NZ, uint size, uint lattice_type_index) {
        gpucad::check(gcb_create_lattice(gpucad::ctx(), d_latticevol, NX, NY, NZ, size, lattice_type_index), "create_lattice");
    }
    // the spectrum part of Multitopo::unit_lattice (main.cu:3577-3706: fft_func + fft_scalar + fft_fill + host pick) in one call:
    // d_lattice_data = device float2[(2*range_st+1)^3], the reference's `lattice_data`
    void unit_lattice_spectrum(const float* d_unit_cell, int Nxu, int Nyu, int Nzu, int range_st, float2* d_lattice_data) {
        gpucad::check(gcb_unit_lattice_spectrum(gpucad::ctx(), d_unit_cell, Nxu, Nyu, Nzu, range_st, d_lattice_data), "unit_lattice_spectrum");
    }
};

class Interpolations {
public:
    void setupTexture(int dx, int dy, int dz) { gpucad::check(gcb_setupTexture(gpucad::ctx(), dx, dy, dz), "setupTexture"); }
    void copytotexture(float* d_phi, cudaPitchedPtr p, int NX, int NY, int NZ) {
        gpucad::check(gcb_copytotexture(gpucad::ctx(), d_phi, gcb_pitched_ptr{p.ptr, p.pitch, p.xsize, p.ysize}, NX, NY, NZ), "copytotexture");
    }
    void updateTexture(cudaPitchedPtr p) { gpucad::check(gcb_updateTexture(gpucad::ctx(), gcb_pitched_ptr{p.ptr, p.pitch, p.xsize, p.ysize}), "updateTexture"); }
    void deleteTexture() { gpucad::check(gcb_deleteTexture(gpucad::ctx()), "deleteTexture"); }
};

class Gratings : public Interpolations {
public:
    int NX = 0, NY = 0, NZ = 0;  // set by the application before the phase solve (main.cu:4182-4186), read by GPUCG_lattice
    // SVL phase solve (Gratings.h:31-41): period / rotation fields, right-hand side of one harmonic, conjugate gradients
    void angle_data(float* d_theta, int NX_, int NY_, int NZ_, float dx, float dy, float dz, float mean_x, float mean_y, float mean_z, char axis) {
        gpucad::check(gcb_angle_data(gpucad::ctx(), d_theta, NX_, NY_, NZ_, dx, dy, dz, mean_x, mean_y, mean_z, axis), "angle_data");
    }
    void period_data(float* d_period, int NX_, int NY_, int NZ_, float dx, float dy, float dz, float mean_x, float mean_y, float mean_z, char axis) {
        gpucad::check(gcb_period_data(gpucad::ctx(), d_period, NX_, NY_, NZ_, dx, dy, dz, mean_x, mean_y, mean_z, axis), "period_data");
    }
    void finding_phi(float* d_phi, float* d_period, int x_dim, int y_dim, int z_dim, int i, int j, int k, float dx, float dy, float dz, char latticetype_one,
                     int unform_type, float const_peirod, float x_period, float y_period, float z_period, float lcon, float lcon_1, bool sinewave_zaxis) {
        gpucad::check(gcb_finding_phi(gpucad::ctx(), d_phi, d_period, x_dim, y_dim, z_dim, i, j, k, dx, dy, dz, latticetype_one, unform_type, const_peirod, x_period,
                                      y_period, z_period, lcon, lcon_1, sinewave_zaxis),
                      "finding_phi");
    }
    void GPUCG_lattice(float* d_phi, const int iter, const int OptIter, const float EndRes, int& FinalIter, float& FinalRes) {
        gpucad::check(gcb_GPUCG_lattice(gpucad::ctx(), d_phi, NX, NY, NZ, iter, OptIter, EndRes, &FinalIter, &FinalRes), "GPUCG_lattice");
    }
    void GPU_buffer_normalise_three(float* dataone, float* datatwo, size_t size, float a1, float b1) {
        gpucad::check(gcb_GPU_buffer_normalise_three(gpucad::ctx(), dataone, datatwo, size, a1, b1), "GPU_buffer_normalise_three");
    }
    void GPU_buffer_normalise_buffer(float* d_vec1, float* d_vec2, int n) { gpucad::check(gcb_GPU_buffer_normalise_buffer(gpucad::ctx(), d_vec1, d_vec2, n), "GPU_buffer_normalise_buffer"); }
    void GPU_buffer_normalise_four(float* dataone, float* datatwo, float* datathree, size_t size, int Nx, int Ny, int Nz, float isoval_1, float isoval_2) {
        gpucad::check(gcb_GPU_buffer_normalise_four(gpucad::ctx(), dataone, datatwo, datathree, size, Nx, Ny, Nz, isoval_1, isoval_2), "GPU_buffer_normalise_four");
    }
    void grating(float2* dvol, int NX2, int NY2, int NZ2, float dx2, float dy2, float dz2) { gpucad::check(gcb_grating(gpucad::ctx(), dvol, NX2, NY2, NZ2, dx2, dy2, dz2), "grating"); }
    void refine(float* dvol, int NX2, int NY2, int NZ2, float dx, float dy, float dz) { gpucad::check(gcb_refine(gpucad::ctx(), dvol, NX2, NY2, NZ2, dx, dy, dz), "refine"); }
    void svl(float* d_svl, float2* d_grating, int NX, int NY, int NZ, int indxx, float2* data_fft) { gpucad::check(gcb_svl(gpucad::ctx(), d_svl, d_grating, NX, NY, NZ, indxx, data_fft), "svl"); }
    void topo_field(float* topo_field, float* isosurf, float volfrac, int NX, int NY, int NZ) { gpucad::check(gcb_topo_field(gpucad::ctx(), topo_field, isosurf, volfrac, NX, NY, NZ), "topo_field"); }
    void primitive_field(grid_points* primitive_field, float* primitive_active, float* isosurf, float isoval, bool fixed, bool active, int NX, int NY, int NZ) {
        gpucad::check(gcb_primitive_field(gpucad::ctx(), (gcb_grid_points*)primitive_field, primitive_active, isosurf, isoval, fixed, active, NX, NY, NZ), "primitive_field");
    }
};

class File_output {
public:
    void file_write_obj(float4* d_pos, uint totalVerts, const char* filename) { gpucad::check(gcb_file_write_obj(gpucad::ctx(), d_pos, totalVerts, filename), "file_write_obj"); }
};
