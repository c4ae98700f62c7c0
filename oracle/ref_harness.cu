/*
 * ref_harness.cu -- headless C-ABI driver around the UNMODIFIED reference translation units.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY.  Compiled by oracle/Makefile together with the
 * reference sources where they lie under /root/reference/src (never copied into this repo)
 * into oracle/_ref/libgpucad_ref.so.  It gives tests/ and `bench.py --impl reference` a way
 * to run the reference's own CUDA kernels on the GPU box, on the same device buffers as the
 * product library, so that stage arrays and meshes can be compared bit for bit.
 *
 * Everything here is a thin call into a reference class method; the call sequences follow
 * src/main.cu (show_model :3304, spatial_lattice_run :3904, display_unit_lattice :4106,
 * toprun_struct :3024).  One deliberate deviation, documented in SURVEY.md appendix A-1:
 * Isosurface::computeIsosurface{,_lattice,_latticeone} drop thread blocks above 65535
 * (`grid.y = grid.x / 32768`, Isosurface.cu:56-60).  `ref_isosurface_*` therefore replays the
 * same host sequence through the MarchingCubeCuda wrappers with a correct 2-D grid when
 * `fix_grid` is non-zero; with fix_grid == 0 the reference method is called as shipped.
 */
#include <cuda_runtime.h>
#include <cufft.h>
#include <cstdio>
#include <cmath>

#include "Isosurface.h"
#include "Modelling.h"
#include "File_output.h"
#include "lattice_files/Fft_lattice.h"
#include "lattice_files/Gratings.h"

/* declared `extern` by lattice_files/Fft_lattice.cu:8-10, defined by main.cu in the app */
cufftHandle planr2c;
cufftHandle planc2r;
cufftHandle planc2c;

namespace {
Isosurface* g_iso = nullptr;
Gratings* g_lat = nullptr;
Fft_lattice g_fft;
File_output g_out;
uint* g_triTable = nullptr;
uint* g_numVertsTable = nullptr;
cudaPitchedPtr g_pitched = {};
bool g_tex = false;

struct GridDesc { uint3 size, shift, mask; uint numVoxels; };
GridDesc make_grid(uint nx, uint ny, uint nz) { /* initMC_two: main.cu:2125-2139 */
    GridDesc g;
    g.size = make_uint3(nx, ny, nz);
    g.mask = make_uint3(nx - 1, ny - 1, nz - 1);
    g.shift = make_uint3(1, nx - 1, (nx - 1) * (ny - 1));
    g.numVoxels = g.mask.x * g.mask.y * g.mask.z;
    return g;
}
dim3 fixed_grid(uint n) {
    uint blocks = (n + 1023u) / 1024u;
    if (blocks > 65535u) return dim3(32768, (blocks + 32767u) / 32768u, 1);
    return dim3(blocks, 1, 1);
}
int last_error(const char* what) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { fprintf(stderr, "ref_harness: %s: %s\n", what, cudaGetErrorString(e)); return 1; }
    return 0;
}
/* scan tail shared by every computeIsosurface* variant (Isosurface.cu:69-116) */
uint scan_total(uint* scan, uint* in, uint n) {
    uint a = 0, b = 0;
    cudaMemcpy(&a, in + n - 1, sizeof(uint), cudaMemcpyDeviceToHost);
    cudaMemcpy(&b, scan + n - 1, sizeof(uint), cudaMemcpyDeviceToHost);
    return a + b;
}
} // namespace

extern "C" {

int ref_init(void) {
    if (!g_iso) {
        g_iso = new Isosurface();
        g_lat = new Gratings();
        g_iso->allocateTextures_s(&g_triTable, &g_numVertsTable); /* init_textures: main.cu:2040-2042 */
    }
    return last_error("ref_init");
}

/* ---- field producers ---- */
int ref_create_lattice(float* d_out, uint nx, uint ny, uint nz, uint type) {
    g_fft.create_lattice(d_out, nx, ny, nz, nx * ny * nz, type);
    return last_error("create_lattice");
}
int ref_sphere(float* d, float cx, float cy, float cz, float r, float t, int nx, int ny, int nz, float dx, float dy, float dz, int shell) {
    Modelling m(nx, ny, nz);
    m.sphere_with_center(d, make_float3(cx, cy, cz), r, t, nx, ny, nz, dx, dy, dz, shell != 0);
    return last_error("sphere");
}
int ref_distance_from_line(float* d, float cx, float cy, float cz, float ax, float ay, float az, float r, float tr, float ta,
                           int nx, int ny, int nz, float dx, float dy, float dz, int disc) {
    Modelling m(nx, ny, nz);
    m.distance_from_line(d, make_float3(cx, cy, cz), make_float3(ax, ay, az), r, tr, ta, nx, ny, nz, dx, dy, dz, disc != 0);
    return last_error("distance_from_line");
}
int ref_cuboid(float* d, float cx, float cy, float cz, float a0, float a1, float a2, float xw, float yw, float zw,
               int nx, int ny, int nz, float dx, float dy, float dz) {
    Modelling m(nx, ny, nz);
    m.cuboid(d, make_float3(cx, cy, cz), make_float3(a0, a1, a2), xw, yw, zw, nx, ny, nz, dx, dy, dz);
    return last_error("cuboid");
}
int ref_cuboid_shell(float* d, float cx, float cy, float cz, float a0, float a1, float a2, float xw, float yw, float zw, float th,
                     int nx, int ny, int nz, float dx, float dy, float dz) {
    Modelling m(nx, ny, nz);
    m.cuboid_shell(d, make_float3(cx, cy, cz), make_float3(a0, a1, a2), xw, yw, zw, th, nx, ny, nz, dx, dy, dz);
    return last_error("cuboid_shell");
}
int ref_torus(float* d, float cx, float cy, float cz, float a0, float a1, float a2, float R, float rc,
              int nx, int ny, int nz, float dx, float dy, float dz) {
    Modelling m(nx, ny, nz);
    m.torus_with_center(d, make_float3(cx, cy, cz), make_float3(a0, a1, a2), R, rc, nx, ny, nz, dx, dy, dz);
    return last_error("torus");
}
int ref_cone(float* d, float cx, float cy, float cz, float a0, float a1, float a2, float br, float h,
             int nx, int ny, int nz, float dx, float dy, float dz) {
    Modelling m(nx, ny, nz);
    m.cone_with_base_radius_height(d, make_float3(cx, cy, cz), make_float3(a0, a1, a2), br, h, nx, ny, nz, dx, dy, dz);
    return last_error("cone");
}
int ref_cone_frustum(float* d, float cx, float cy, float cz, float a0, float a1, float a2, float tr, float br, float h,
                     int nx, int ny, int nz, float dx, float dy, float dz) {
    Modelling m(nx, ny, nz);
    m.cone_frustum(d, make_float3(cx, cy, cz), make_float3(a0, a1, a2), tr, br, h, nx, ny, nz, dx, dy, dz);
    return last_error("cone_frustum");
}
int ref_pyramid_frustum(float* d, float cx, float cy, float cz, float a0, float a1, float a2, float xb, float xt, float yh,
                        float zb, float zt, int nx, int ny, int nz, float dx, float dy, float dz) {
    Modelling m(nx, ny, nz);
    m.pyramid_frustum(d, make_float3(cx, cy, cz), make_float3(a0, a1, a2), xb, xt, yh, zb, zt, nx, ny, nz, dx, dy, dz);
    return last_error("pyramid_frustum");
}

int ref_normalise_buffer(float* d_in, float* d_out, int n) {
    g_lat->GPU_buffer_normalise_buffer(d_in, d_out, n);
    return last_error("normalise_buffer");
}
int ref_normalise_four(float* d_in, float* d_mask, float* d_k, size_t size, int nx, int ny, int nz, float iso1, float iso2) {
    g_lat->GPU_buffer_normalise_four(d_in, d_mask, d_k, size, nx, ny, nz, iso1, iso2);
    return last_error("normalise_four");
}
int ref_primitive_field(grid_points* prim, float* active, float* isosurf, int fixed, int dynamic, int nx, int ny, int nz) {
    g_lat->primitive_field(prim, active, isosurf, 0.0f, fixed != 0, dynamic != 0, nx, ny, nz);
    return last_error("primitive_field");
}
int ref_topo_field(float* topo, float* isosurf, float volfrac, int nx, int ny, int nz) {
    g_lat->topo_field(topo, isosurf, volfrac, nx, ny, nz);
    return last_error("topo_field");
}
int ref_patch_topo_field(float* d, int nx, int ny, int nz, grid_points* vol_one) {
    g_iso->patch_topo_field(d, nx, ny, nz, vol_one);
    return last_error("patch_topo_field");
}

/* ---- control grid texture: init_textures main.cu:2046-2056, spatial_lattice_run :3964-3970 ---- */
int ref_setup_texture(int cx, int cy, int cz) {
    if (g_tex) { g_lat->deleteTexture(); cudaFree(g_pitched.ptr); g_tex = false; }
    cudaExtent ext = make_cudaExtent((size_t)cx * sizeof(float), cy, cz);
    if (cudaMalloc3D(&g_pitched, ext) != cudaSuccess) return 1;
    g_lat->setupTexture(cx, cy, cz);
    g_tex = true;
    return last_error("setup_texture");
}
int ref_upload_texture(float* d_phi, int cx, int cy, int cz) {
    g_lat->copytotexture(d_phi, g_pitched, cx, cy, cz);
    g_lat->updateTexture(g_pitched);
    return last_error("upload_texture");
}
int ref_delete_texture(void) {
    if (g_tex) { g_lat->deleteTexture(); cudaFree(g_pitched.ptr); g_tex = false; }
    return last_error("delete_texture");
}
int ref_refine(float* d_out, int nx2, int ny2, int nz2, float dx, float dy, float dz) {
    g_lat->refine(d_out, nx2, ny2, nz2, dx, dy, dz);
    return last_error("refine");
}
int ref_grating(float2* d_ga, int nx2, int ny2, int nz2, float dx, float dy, float dz) {
    g_lat->grating(d_ga, nx2, ny2, nz2, dx, dy, dz);
    return last_error("grating");
}
int ref_svl(float* d_svl, float2* d_ga, int nx, int ny, int nz, int idx, float2* d_coef) {
    g_lat->svl(d_svl, d_ga, nx, ny, nz, idx, d_coef);
    return last_error("svl");
}
/* whole SVL field as spatial_lattice_run accumulates it (without the per-harmonic redraw):
 * for h: copytotexture, updateTexture, grating, svl.  d_phi holds nh control grids back to back. */
int ref_svl_field(float* d_svl, float2* d_ga, float* d_phi, int nh, float2* d_coef, int cx, int cy, int cz,
                  int nx2, int ny2, int nz2, float dx, float dy, float dz) {
    for (int h = 0; h < nh; ++h) {
        g_lat->copytotexture(d_phi + (size_t)h * cx * cy * cz, g_pitched, cx, cy, cz);
        g_lat->updateTexture(g_pitched);
        g_lat->grating(d_ga, nx2, ny2, nz2, dx, dy, dz);
        g_lat->svl(d_svl, d_ga, nx2, ny2, nz2, h, d_coef);
    }
    return last_error("svl_field");
}

/* ---- CSG retain ---- */
int ref_copy_parameter(uint* voxel_verts, float iso, uint nx, uint ny, uint nz, float vx, float vy, float vz,
                       grid_points* vol_one, float* vol_two, float* vol_lattice, int fixed, int dynamic, float iso1, float iso2,
                       int obj_union, int obj_diff, int obj_intersect) {
    GridDesc g = make_grid(nx, ny, nz);
    g_iso->copy_parameter(voxel_verts, iso, g.size, g.shift, g.mask, make_float3(vx, vy, vz), g.numVoxels, vol_one, vol_two,
                          vol_lattice, fixed != 0, dynamic != 0, iso1, iso2, obj_union != 0, obj_diff != 0, obj_intersect != 0);
    return last_error("copy_parameter");
}

/* ---- extraction ---- */
struct RefScratch { uint *verts, *vertsScan, *occ, *occScan, *comp; };

int ref_isosurface_lattice(int one, int fix_grid, float* vol, float4* pos, float4* norm, float iso, uint nx, uint ny, uint nz,
                           float vx, float vy, float vz, float gx, float gy, float gz, uint* d_verts, uint* d_vertsScan,
                           uint* d_occ, uint* d_occScan, uint* d_comp, uint maxVerts, float* vol_one, float* vol_two,
                           float isovalue1, float isovalue2, float iso1, float iso2, uint* activeVoxels, uint* totalVerts) {
    GridDesc g = make_grid(nx, ny, nz);
    float3 vs = make_float3(vx, vy, vz), gc = make_float3(gx, gy, gz);
    if (!fix_grid) {
        if (one)
            g_iso->computeIsosurface_latticeone(vol, pos, norm, iso, g.numVoxels, d_verts, d_vertsScan, d_occ, d_occScan, g.size,
                                                g.shift, g.mask, vs, gc, activeVoxels, totalVerts, d_comp, maxVerts, vol_one,
                                                isovalue1, isovalue2);
        else
            g_iso->computeIsosurface_lattice(vol, pos, norm, iso, g.numVoxels, d_verts, d_vertsScan, d_occ, d_occScan, g.size,
                                             g.shift, g.mask, vs, gc, activeVoxels, totalVerts, d_comp, maxVerts, vol_one, vol_two,
                                             isovalue1, isovalue2, iso1, iso2);
        return last_error("computeIsosurface_lattice*");
    }
    /* Isosurface.cu:401-572 replayed with a grid that covers every cell */
    dim3 threads(1024, 1, 1);
    g_iso->classifyVoxel_lattice_new(fixed_grid(g.numVoxels), threads, d_verts, d_occ, vol, g.size, g.shift, g.mask, g.numVoxels, vs, iso);
    g_iso->ThrustScanWrapper_lattice(d_occScan, d_occ, g.numVoxels);
    *activeVoxels = scan_total(d_occScan, d_occ, g.numVoxels);
    if (*activeVoxels == 0) { *totalVerts = 0; return last_error("lattice(empty)"); }
    g_iso->compactVoxels_lattice(fixed_grid(g.numVoxels), threads, d_comp, d_occ, d_occScan, g.numVoxels);
    g_iso->ThrustScanWrapper_lattice(d_vertsScan, d_verts, g.numVoxels);
    *totalVerts = scan_total(d_vertsScan, d_verts, g.numVoxels);
    dim3 grid2((*activeVoxels + NTHREADS - 1) / NTHREADS, 1, 1), tids2(NTHREADS, 1, 1);
    if (one)
        g_iso->generateTriangles_lattice_newone(grid2, tids2, pos, norm, d_comp, d_vertsScan, vol, g.size, g.shift, g.mask, vs, gc, iso,
                                                *activeVoxels, maxVerts, *totalVerts, vol_one, isovalue1, isovalue2);
    else
        g_iso->generateTriangles_lattice_new(grid2, tids2, pos, norm, d_comp, d_vertsScan, vol, g.size, g.shift, g.mask, vs, gc, iso,
                                             *activeVoxels, maxVerts, *totalVerts, vol_one, vol_two, isovalue1, isovalue2, iso1, iso2);
    return last_error("lattice(fixed grid)");
}

int ref_isosurface_csg(int fix_grid, float4* pos, float4* norm, float iso, uint nx, uint ny, uint nz, float vx, float vy, float vz,
                       float gx, float gy, float gz, uint* d_verts, uint* d_vertsScan, uint* d_occ, uint* d_occScan, uint* d_comp,
                       uint maxVerts, grid_points* fixed_f, float* dynamic_f, float* topo_f, float* lattice_f, float iso1, float iso2,
                       int obj_union, int obj_diff, int obj_intersect, int fixed, int dynamic, int make_region,
                       uint* activeVoxels, uint* totalVerts) {
    GridDesc g = make_grid(nx, ny, nz);
    float3 vs = make_float3(vx, vy, vz), gc = make_float3(gx, gy, gz);
    size_t nfacets = 0;
    bool U = obj_union != 0, D = obj_diff != 0, I = obj_intersect != 0, F = fixed != 0, Y = dynamic != 0, M = make_region != 0;
    if (!fix_grid) {
        g_iso->computeIsosurface(nullptr, g.size, pos, norm, iso, g.numVoxels, d_verts, d_vertsScan, d_occ, d_occScan, g.size, g.shift,
                                 g.mask, vs, gc, activeVoxels, totalVerts, d_comp, maxVerts, fixed_f, dynamic_f, topo_f, lattice_f, iso1,
                                 iso2, U, D, I, true, false, false, F, Y, M, &nfacets);
        return last_error("computeIsosurface");
    }
    dim3 threads(1024, 1, 1);
    g_iso->classifyVoxel_lattice(fixed_grid(g.numVoxels), threads, nullptr, g.size, d_verts, d_occ, fixed_f, dynamic_f, topo_f, lattice_f,
                                 g.size, g.shift, g.mask, g.numVoxels, iso1, iso2, vs, iso, U, D, I, true, false, false, F, Y, M);
    g_iso->ThrustScanWrapper_lattice(d_occScan, d_occ, g.numVoxels);
    *activeVoxels = scan_total(d_occScan, d_occ, g.numVoxels);
    if (*activeVoxels == 0) { *totalVerts = 0; return last_error("csg(empty)"); }
    g_iso->compactVoxels_lattice(fixed_grid(g.numVoxels), threads, d_comp, d_occ, d_occScan, g.numVoxels);
    g_iso->ThrustScanWrapper_lattice(d_vertsScan, d_verts, g.numVoxels);
    *totalVerts = scan_total(d_vertsScan, d_verts, g.numVoxels);
    dim3 grid2((*activeVoxels + NTHREADS - 1) / NTHREADS, 1, 1), tids2(NTHREADS, 1, 1);
    g_iso->generateTriangles_lattice(grid2, tids2, pos, norm, d_comp, d_vertsScan, g.size, g.shift, g.mask, vs, gc, iso, *activeVoxels,
                                     maxVerts, *totalVerts, fixed_f, dynamic_f, topo_f, lattice_f, iso1, iso2, d_verts, U, D, I, true,
                                     false, false, F, Y, M);
    return last_error("csg(fixed grid)");
}

/* Isosurface::computeIsosurface_region as shipped (Isosurface.cu:150-239); fix_grid replays it with a grid covering every cell */
int ref_isosurface_region(int fix_grid, float4* pos, float4* norm, float iso, uint nx, uint ny, uint nz, float vx, float vy, float vz,
                          float gx, float gy, float gz, uint* d_verts, uint* d_vertsScan, uint* d_occ, uint* d_occScan, uint* d_comp,
                          uint maxVerts, grid_points* vol_topo, grid_points* fixed_f, float* dynamic_f, int make_region, int show_region,
                          int show_domain, triangle_metadata* triangle_data, uint* activeVoxels, uint* totalVerts) {
    GridDesc g = make_grid(nx, ny, nz);
    float3 vs = make_float3(vx, vy, vz), gc = make_float3(gx, gy, gz);
    bool M = make_region != 0, R = show_region != 0, D = show_domain != 0;
    if (!fix_grid) {
        g_iso->computeIsosurface_region(pos, norm, iso, g.numVoxels, d_verts, d_vertsScan, d_occ, d_occScan, g.size, g.shift, g.mask, vs, gc,
                                        activeVoxels, totalVerts, d_comp, maxVerts, vol_topo, fixed_f, dynamic_f, nullptr, nullptr, 0.f, 0.f,
                                        true, false, false, true, false, false, false, false, M, R, D, triangle_data);
        return last_error("computeIsosurface_region");
    }
    dim3 threads(1024, 1, 1);
    g_iso->classifyVoxel_region(fixed_grid(g.numVoxels), threads, d_verts, d_occ, vol_topo, fixed_f, dynamic_f, nullptr, nullptr, g.size, g.shift,
                                g.mask, g.numVoxels, 0.f, 0.f, vs, iso, true, false, false, true, false, false, false, false, M, R, D);
    g_iso->ThrustScanWrapper_lattice(d_occScan, d_occ, g.numVoxels);
    *activeVoxels = scan_total(d_occScan, d_occ, g.numVoxels);
    if (*activeVoxels == 0) { *totalVerts = 0; return last_error("region(empty)"); }
    g_iso->compactVoxels_lattice(fixed_grid(g.numVoxels), threads, d_comp, d_occ, d_occScan, g.numVoxels);
    g_iso->ThrustScanWrapper_lattice(d_vertsScan, d_verts, g.numVoxels);
    *totalVerts = scan_total(d_vertsScan, d_verts, g.numVoxels);
    cudaMemset(pos, 0, maxVerts);
    cudaMemset(norm, 0, maxVerts);
    dim3 grid2((*activeVoxels + NTHREADS - 1) / NTHREADS, 1, 1), tids2(NTHREADS, 1, 1);
    g_iso->generateTriangles_region(grid2, tids2, pos, norm, d_comp, d_vertsScan, g.size, g.shift, g.mask, vs, gc, iso, *activeVoxels, maxVerts,
                                    *totalVerts, vol_topo, fixed_f, dynamic_f, nullptr, nullptr, 0.f, 0.f, d_verts, true, false, false, true, false,
                                    false, false, false, M, R, D, triangle_data);
    return last_error("region(fixed grid)");
}

int ref_isosurface_topo(int with_disp_variant, float4* pos, float4* norm, float iso, uint nx, uint ny, uint nz, float vx, float vy,
                        float vz, float gx, float gy, float gz, uint* d_verts, uint* d_vertsScan, uint* d_occ, uint* d_occScan,
                        uint* d_comp, uint maxVerts, grid_points* vol_topo, grid_points* vol_one, float* vol_two, float* d_solid,
                        float isovalue1, float* d_result, int disp, float4* disp_two, uint* activeVoxels, uint* totalVerts) {
    GridDesc g = make_grid(nx, ny, nz);
    float3 vs = make_float3(vx, vy, vz), gc = make_float3(gx, gy, gz);
    if (with_disp_variant)
        g_iso->computeIsosurface_topo(pos, norm, iso, g.numVoxels, d_verts, d_vertsScan, d_occ, d_occScan, g.size, g.shift, g.mask, vs, gc,
                                      activeVoxels, totalVerts, d_comp, maxVerts, vol_topo, vol_one, vol_two, d_solid, isovalue1,
                                      d_result, nullptr, disp != 0, disp_two);
    else
        g_iso->computeIsosurface_2(pos, norm, iso, g.numVoxels, d_verts, d_vertsScan, d_occ, d_occScan, g.size, g.shift, g.mask, vs, gc,
                                   activeVoxels, totalVerts, d_comp, maxVerts, vol_topo, vol_one, vol_two, d_solid, isovalue1, d_result,
                                   nullptr);
    return last_error("computeIsosurface_topo/_2");
}

int ref_write_obj(float4* d_pos, uint totalVerts, const char* filename) {
    g_out.file_write_obj(d_pos, totalVerts, filename);
    return last_error("file_write_obj");
}


/* ---- SVL phase solve: Gratings::period_data + GPU_buffer_normalise_three, finding_phi, GPUCG_lattice as
 * Multitopo::spatial_lattice_run calls them (main.cu:3927-3962) */
int ref_period_data(float* d_period, int nx, int ny, int nz, float dx, float dy, float dz, float mx, float my, float mz, int axis) {
    g_lat->period_data(d_period, nx, ny, nz, dx, dy, dz, mx, my, mz, (char)axis);
    return last_error("period_data");
}
int ref_angle_data(float* d_theta, int nx, int ny, int nz, float dx, float dy, float dz, float mx, float my, float mz, int axis) {
    g_lat->angle_data(d_theta, nx, ny, nz, dx, dy, dz, mx, my, mz, (char)axis);
    return last_error("angle_data");
}
int ref_normalise_three(float* d_in, float* d_out, size_t size, float a1, float b1) {
    g_lat->GPU_buffer_normalise_three(d_in, d_out, size, a1, b1);
    return last_error("normalise_three");
}
int ref_finding_phi(float* d_phi, float* d_period, int nx, int ny, int nz, int i, int j, int k, float dx, float dy, float dz, int latticetype, int uniform_type,
                    float const_period, float x_period, float y_period, float z_period, float lcon, float lcon_1, int sinewave_zaxis) {
    g_lat->finding_phi(d_phi, d_period, nx, ny, nz, i, j, k, dx, dy, dz, (char)latticetype, uniform_type, const_period, x_period, y_period, z_period, lcon, lcon_1,
                       sinewave_zaxis != 0);
    return last_error("finding_phi");
}
int ref_cg(float* d_phi, int nx, int ny, int nz, int iter, float end_res, int* final_iter, float* final_res) {
    g_lat->NX = nx; g_lat->NY = ny; g_lat->NZ = nz;   /* main.cu:4182-4184 */
    int fi = 0; float fr = 0.f;
    g_lat->GPUCG_lattice(d_phi, iter, 1, end_res, fi, fr);
    *final_iter = fi; *final_res = fr;
    return last_error("GPUCG_lattice");
}
} // extern "C"
