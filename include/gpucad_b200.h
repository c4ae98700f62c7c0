/*
 * gpucad_b200.h -- C ABI of the B200-native implicit-field + marching-cubes engine.
 *
 * Drop-in boundary for ONE path of DESIGN4ADDITIVE/GPUCADforAM: implicit field on a voxel grid
 * -> classify -> scan -> compact -> triangles + normals (-> .obj).  The reference has no FFI
 * layer; its boundary is the set of C++ member functions `Multitopo` (src/main.cu) calls with
 * raw device pointers.  Every "legacy" entry point below keeps the parameter ORDER, buffer
 * LAYOUTS and result semantics of the reference method it replaces (cited per function), with
 * CUDA vector types replaced by layout-identical PODs and `bool` by int.  All pointers named
 * d_* / documented "device" are CUDA device pointers owned by the caller; the library never
 * frees or reallocates them.  Results the reference returns through host pointers
 * (activeVoxels, totalVerts, nfacets) are written to host pointers here as well.
 *
 * Errors: every function returns 0 on success, non-zero on failure (gcb_last_error() gives the
 * text).  The reference prints and exit(1)s (commons/helper_cuda.h:583-612); a C++ shim that
 * wants that behaviour wraps the return code.
 *
 * There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef GPUCAD_B200_H
#define GPUCAD_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* layout-identical to CUDA's uint3 / float3 / float2 / float4 / cudaPitchedPtr */
typedef struct gcb_uint3 { unsigned int x, y, z; } gcb_uint3;
typedef struct gcb_float3 { float x, y, z; } gcb_float3;
typedef struct gcb_pitched_ptr { void* ptr; size_t pitch, xsize, ysize; } gcb_pitched_ptr;
/* reference: struct grid_points, src/MarchingCubes_kernel.h:12-18 (16-byte AoS) */
typedef struct gcb_grid_points { int val; float t_x, t_y, t_z; } gcb_grid_points;

/* reference: struct triangle_metadata, src/MarchingCubes_kernel.h:20-32 (64 bytes, 4-byte aligned).  computeIsosurface_region
 * writes index .. edge_3, centroid and normal; load_group and force_dir belong to the picking code and are left untouched. */
typedef struct gcb_triangle_metadata {
    unsigned int index, voxel, l_index, edge_1, edge_2, edge_3, load_group;
    float centroid[3], normal[3], force_dir[3];
} gcb_triangle_metadata;

typedef struct gcb_ctx gcb_ctx;

/* ------------------------------------------------------------------ context */
/* stream: a cudaStream_t (NULL = legacy default stream, as the reference uses). */
int gcb_create(gcb_ctx** ctx, int device, void* stream);
int gcb_destroy(gcb_ctx* ctx);
const char* gcb_last_error(gcb_ctx* ctx);
int gcb_set_stream(gcb_ctx* ctx, void* stream);
/* option flags */
enum {
    GCB_OPT_FILL_STAGE_ARRAYS = 1, /* also write d_voxelVerts/_Scan/d_voxelOccupied/_Scan (parity tests) */
    GCB_OPT_LEGACY_MEMSET = 2,     /* cudaMemset(pos/norm, 0, maxVerts BYTES) as Isosurface.cu:120-121 (default on).  One divergence:
                                      the reference clears only after it knows activeVoxels > 0 and leaves both buffers untouched for
                                      an empty surface; here the clear is enqueued before the fused kernel, so it also happens when
                                      activeVoxels == 0 (consumers honour totalVerts either way) */
    GCB_OPT_NO_TMA = 4,            /* force the LDG stage-in path (debug / A-B measurement) */
    GCB_OPT_OBJ_HOST = 8,          /* gcb_file_write_obj: weld and format on one host thread (as the reference does) instead of on the GPU */
    GCB_OPT_FAST_FIELD = 32,       /* gcb_svl_field / gcb_svl_lattice*: evaluate the SVL sum as |c_h| cos(phi_h + arg c_h) with the hardware
                                      cosine and packed-fp32 trilinear blends instead of the bit-exact texture model + libdevice sincosf.
                                      The field then agrees with the default mode to ~1e-6 of its range (bound: sum_h |c_h| (ulp(phi_h)/2 +
                                      4e-6), the first term being the rounding the reference itself applies to the interpolated phase); the
                                      extraction is unchanged, i.e. bit-exact on that field.  Power-of-two upsampling ratios only (others
                                      fall back to the exact kernels); valid for |phi| < 2^24.  Off by default */
    GCB_OPT_ASYNC_FIELDS = 16      /* legacy calls without host results (primitives, create_lattice, normalise, refine, grating, svl, copy_parameter,
                                      texture upload ...) only enqueue on the context's stream instead of ending in a device synchronise as the
                                      reference wrappers do (MarchingCubes_kernel.cu:458, :1074); calls that report counts still synchronise.  Off by
                                      default: the default keeps the reference's blocking semantics */
};
int gcb_set_options(gcb_ctx* ctx, unsigned int flags);
/* number of kernels this library launched on the context since creation / last reset */
unsigned long long gcb_launch_count(gcb_ctx* ctx);
void gcb_reset_launch_count(gcb_ctx* ctx);
/* device time (ms) of the extraction kernel of the last extraction call, measured with CUDA
 * events on the context's stream; < 0 if timing was not enabled */
int gcb_enable_kernel_timing(gcb_ctx* ctx, int on);
float gcb_last_extract_kernel_ms(gcb_ctx* ctx);
float gcb_last_field_kernel_ms(gcb_ctx* ctx);

/* MarchingCubeCuda::allocateTextures_s / destroyAllTextureObjects (MarchingCubes_kernel.cu:24-75).
 * The tables live in __constant__ memory of this library; these exist for call-sequence parity
 * and additionally hand back device copies in the reference's uint32 layout. */
int gcb_allocateTextures_s(gcb_ctx* ctx, unsigned int** d_triTable, unsigned int** d_numVertsTable);
int gcb_destroyAllTextureObjects(gcb_ctx* ctx);
/* host copies of the tables: tri[256*16] (255 = end), nverts[256] */
void gcb_tables(unsigned int* tri, unsigned int* nverts);

/* ------------------------------------------------------------------ extraction (legacy signatures) */
/* Capacity rule shared by every extraction entry point: maxVerts is the capacity of pos / norm in VERTICES and a triangle is
 * written iff its first vertex index satisfies `index < maxVerts - 3` (the reference's guard, MarchingCubes_kernel.cu:2181) --
 * so a buffer that must hold all totalVerts vertices needs maxVerts >= totalVerts + 3, and with maxVerts == totalVerts the last
 * triangle is dropped exactly as in the reference.  maxVerts < 3 writes nothing (the reference's unsigned `maxVerts - 3` wraps
 * there and writes out of bounds; that is not reproduced).  gcb_extract_band_raw / gcb_svl_lattice* with maxVerts > 2^32 - 1
 * use the exact test `index + 3 <= maxVerts`.  Counts (activeVoxels, totalVerts) are always the untruncated ones. */
/* Isosurface::computeIsosurface  (src/Isosurface.h:28-33, Isosurface.cu:44-134) -- CSG / primitives */
int gcb_computeIsosurface(gcb_ctx* ctx, float* vol, gcb_uint3 raster_grid, void* pos, void* norm, float isoValue,
    unsigned int numVoxels, unsigned int* d_voxelVerts, unsigned int* d_voxelVertsScan, unsigned int* d_voxelOccupied,
    unsigned int* d_voxelOccupiedScan, gcb_uint3 gridSize, gcb_uint3 gridSizeShift, gcb_uint3 gridSizeMask,
    gcb_float3 voxelSize, gcb_float3 gridcenter, unsigned int* activeVoxels, unsigned int* totalVerts,
    unsigned int* d_compVoxelArray, unsigned int maxVerts, gcb_grid_points* primitive_fixed, float* primitive_dynamic,
    float* topo_field, float* lattice_field, float iso1, float iso2, int obj_union, int obj_diff, int obj_intersect,
    int primitive, int topo, int compute_lattice, int fixed, int dynamic, int make_region, size_t* nfacets);

/* Isosurface::computeIsosurface_region  (src/Isosurface.h:38-43, Isosurface.cu:150-239; kernels MarchingCubes_kernel.cu:1163-1302,
 * :2222-2605) -- the GUI's region / domain display and load-support picking.  Cascade per cell: vol_topo (norm.w = 1), then
 * primitive_fixed [& primitive_dynamic for make_region] (0.25), then primitive_fixed (0.5); normals are normalised; with
 * show_region one gcb_triangle_metadata record per triangle (index, voxel, l_index, edge_1..3, centroid, normal) is written to
 * triangle_data, which must hold totalVerts/3 records.  Flag precedence show_region > show_domain > make_region as in the
 * reference; with none set the reference reads an uninitialised cube index, this entry point returns an error instead. */
int gcb_computeIsosurface_region(gcb_ctx* ctx, void* pos, void* norm, float isoValue, unsigned int numVoxels, unsigned int* d_voxelVerts,
    unsigned int* d_voxelVertsScan, unsigned int* d_voxelOccupied, unsigned int* d_voxelOccupiedScan, gcb_uint3 gridSize,
    gcb_uint3 gridSizeShift, gcb_uint3 gridSizeMask, gcb_float3 voxelSize, gcb_float3 gridcenter, unsigned int* activeVoxels,
    unsigned int* totalVerts, unsigned int* d_compVoxelArray, unsigned int maxVerts, gcb_grid_points* vol_topo,
    gcb_grid_points* primitive_fixed, float* primitive_dynamic, float* topo_field, float* lattice_field, float iso1, float iso2,
    int obj_union, int obj_diff, int obj_intersect, int primitive, int topo, int compute_lattice, int fixed, int dynamic,
    int make_region, int show_region, int show_domain, gcb_triangle_metadata* triangle_data);

/* Isosurface::computeIsosurface_lattice  (Isosurface.h:60-64, Isosurface.cu:401-486) */
int gcb_computeIsosurface_lattice(gcb_ctx* ctx, float* vol, void* pos, void* norm, float isoValue, unsigned int numVoxels,
    unsigned int* d_voxelVerts, unsigned int* d_voxelVertsScan, unsigned int* d_voxelOccupied, unsigned int* d_voxelOccupiedScan,
    gcb_uint3 gridSize, gcb_uint3 gridSizeShift, gcb_uint3 gridSizeMask, gcb_float3 voxelSize, gcb_float3 gridcenter,
    unsigned int* activeVoxels, unsigned int* totalVerts, unsigned int* d_compVoxelArray, unsigned int maxVerts,
    float* vol_one, float* vol_two, float isovalue1, float isovalue2, float iso1, float iso2);

/* Isosurface::computeIsosurface_latticeone  (Isosurface.h:66-70, Isosurface.cu:488-572) */
int gcb_computeIsosurface_latticeone(gcb_ctx* ctx, float* vol, void* pos, void* norm, float isoValue, unsigned int numVoxels,
    unsigned int* d_voxelVerts, unsigned int* d_voxelVertsScan, unsigned int* d_voxelOccupied, unsigned int* d_voxelOccupiedScan,
    gcb_uint3 gridSize, gcb_uint3 gridSizeShift, gcb_uint3 gridSizeMask, gcb_float3 voxelSize, gcb_float3 gridcenter,
    unsigned int* activeVoxels, unsigned int* totalVerts, unsigned int* d_compVoxelArray, unsigned int maxVerts,
    float* vol_one, float isovalue1, float isovalue2);

/* Isosurface::computeIsosurface_2  (Isosurface.h:46-50, Isosurface.cu:323-398) */
int gcb_computeIsosurface_2(gcb_ctx* ctx, void* pos, void* norm, float isoValue, unsigned int numVoxels,
    unsigned int* d_voxelVerts, unsigned int* d_voxelVertsScan, unsigned int* d_voxelOccupied, unsigned int* d_voxelOccupiedScan,
    gcb_uint3 gridSize, gcb_uint3 gridSizeShift, gcb_uint3 gridSizeMask, gcb_float3 voxelSize, gcb_float3 gridcenter,
    unsigned int* activeVoxels, unsigned int* totalVerts, unsigned int* d_compVoxelArray, unsigned int maxVerts,
    gcb_grid_points* vol_topo, gcb_grid_points* vol_one, float* vol_two, float* d_solid, float isovalue1, float* d_result,
    void* triangle_data);

/* Isosurface::computeIsosurface_topo  (Isosurface.h:53-58, Isosurface.cu:243-320) */
int gcb_computeIsosurface_topo(gcb_ctx* ctx, void* pos, void* norm, float isoValue, unsigned int numVoxels,
    unsigned int* d_voxelVerts, unsigned int* d_voxelVertsScan, unsigned int* d_voxelOccupied, unsigned int* d_voxelOccupiedScan,
    gcb_uint3 gridSize, gcb_uint3 gridSizeShift, gcb_uint3 gridSizeMask, gcb_float3 voxelSize, gcb_float3 gridcenter,
    unsigned int* activeVoxels, unsigned int* totalVerts, unsigned int* d_compVoxelArray, unsigned int maxVerts,
    gcb_grid_points* vol_topo, gcb_grid_points* vol_one, float* vol_two, float* d_solid, float isovalue1, float* d_result,
    void* triangle_data, int disp, void* disp_two);

/* Isosurface::copy_parameter  (Isosurface.h:19-21, Isosurface.cu:15-27; kernel MarchingCubes_kernel.cu:158-447) */
int gcb_copy_parameter(gcb_ctx* ctx, unsigned int* voxel_verts, float isoValue, gcb_uint3 gridSize, gcb_uint3 gridSizeShift,
    gcb_uint3 gridSizeMask, gcb_float3 voxelSize, unsigned int numVoxels, gcb_grid_points* vol_one, float* vol_two,
    float* vol_lattice, int fixed, int dynamic, float iso1, float iso2, int obj_union, int obj_diff, int obj_intersect);

/* Isosurface::patch_topo_field  (Isosurface.h:74, Isosurface.cu:674-722) */
int gcb_patch_topo_field(gcb_ctx* ctx, float* d_vec1, int Nx, int Ny, int Nz, gcb_grid_points* vol_one);

/* ------------------------------------------------------------------ field producers (legacy signatures) */
/* Modelling::* (src/Modelling.h:26-40, kernels Modelling.cu:244-750) */
int gcb_distance_from_line(gcb_ctx* ctx, float* data_1, gcb_float3 center, gcb_float3 axis, float radius_1, float thickness_radial,
    float thickness_axial, int Nx, int Ny, int Nz, float dx, float dy, float dz, int cylind_disc_selected);
int gcb_sphere_with_center(gcb_ctx* ctx, float* data_1, gcb_float3 center, float radius_1, float thickness_wall, int Nx, int Ny, int Nz,
    float dx, float dy, float dz, int sphere_shell_selected);
int gcb_cuboid(gcb_ctx* ctx, float* data_1, gcb_float3 center, gcb_float3 angles, float x_width, float y_width, float z_width,
    int Nx, int Ny, int Nz, float dx, float dy, float dz);
int gcb_cuboid_shell(gcb_ctx* ctx, float* data_1, gcb_float3 center, gcb_float3 angles, float x_width, float y_width, float z_width,
    float thickness, int Nx, int Ny, int Nz, float dx, float dy, float dz);
int gcb_torus_with_center(gcb_ctx* ctx, float* data_1, gcb_float3 center, gcb_float3 angles, float torus_radius,
    float torus_circle_radius, int Nx, int Ny, int Nz, float dx, float dy, float dz);
int gcb_cone_with_base_radius_height(gcb_ctx* ctx, float* data_1, gcb_float3 center, gcb_float3 angles, float base_radius,
    float cone_height, int Nx, int Ny, int Nz, float dx, float dy, float dz);
int gcb_cone_frustum(gcb_ctx* ctx, float* data_1, gcb_float3 center, gcb_float3 angles, float top_radius, float bottom_radius,
    float cone_frustum_height, int Nx, int Ny, int Nz, float dx, float dy, float dz);
int gcb_pyramid_frustum(gcb_ctx* ctx, float* data_1, gcb_float3 center, gcb_float3 angles, float x_width_base, float x_width_top,
    float y_height, float z_width_base, float z_width_top, int Nx, int Ny, int Nz, float dx, float dy, float dz);

/* Fft_lattice::create_lattice (lattice_files/Fft_lattice.h:15, Fft_lattice.cu:12-74) */
int gcb_create_lattice(gcb_ctx* ctx, float* d_latticevol, unsigned int NX, unsigned int NY, unsigned int NZ, unsigned int size,
    unsigned int lattice_type_index);

/* Gratings::* (lattice_files/Gratings.h:46-77) */
int gcb_GPU_buffer_normalise_buffer(gcb_ctx* ctx, float* d_vec1, float* d_vec2, int n);               /* Gratings.cu:1500-1537 */
int gcb_GPU_buffer_normalise_four(gcb_ctx* ctx, float* dataone, float* datatwo, float* datathree, size_t size, int Nx, int Ny,
    int Nz, float isoval_1, float isoval_2);                                                            /* Gratings.cu:1579-1617 */
/* Gratings::GPU_buffer_normalise_three (Gratings.h:48, Gratings.cu:1539-1572; kernel device_bufferthree :1071-1087):
 * datatwo = a1 + b1 * (dataone - min) / (max - min), min / max with the reduction's clamp through zero. */
int gcb_GPU_buffer_normalise_three(gcb_ctx* ctx, float* dataone, float* datatwo, size_t size, float a1, float b1);
/* Gratings::period_data / angle_data (Gratings.h:31-35, Gratings.cu:1357-1392; kernels :775-853): the spatially varying period
 * (distance from the shifted axis, before GPU_buffer_normalise_three maps it to [a1, a1 + b1]) and rotation fields of the SVL
 * phase solve, Multitopo::spatial_lattice_run main.cu:3927-3931.  axis: 'x', 'y' or 'z'. */
int gcb_period_data(gcb_ctx* ctx, float* d_period, int NX, int NY, int NZ, float dx, float dy, float dz, float mean_x, float mean_y, float mean_z, char axis);
int gcb_angle_data(gcb_ctx* ctx, float* d_theta, int NX, int NY, int NZ, float dx, float dy, float dz, float mean_x, float mean_y, float mean_z, char axis);
int gcb_grating(gcb_ctx* ctx, void* dvol /*float2*/, int NX2, int NY2, int NZ2, float dx2, float dy2, float dz2); /* :976-982 */
int gcb_refine(gcb_ctx* ctx, float* dvol, int NX2, int NY2, int NZ2, float dx, float dy, float dz);   /* :984-990 */
int gcb_svl(gcb_ctx* ctx, float* d_svl, void* d_grating /*float2*/, int NX, int NY, int NZ, int indxx, void* data_fft /*float2*/); /* :992-998 */
int gcb_topo_field(gcb_ctx* ctx, float* topo_field, float* isosurf, float volfrac, int NX, int NY, int NZ); /* :1685-1692 */
int gcb_primitive_field(gcb_ctx* ctx, gcb_grid_points* primitive_field, float* primitive_active, float* isosurf, float isoval,
    int fixed, int active, int NX, int NY, int NZ);                                                     /* :1727-1735 */

/* Interpolations::* (src/Interpolations.h:18-32, Interpolations.cu:16-107).  The "texture" is a
 * context-owned linear control grid sampled by a software trilinear fetch with the texture unit's
 * addressing rules (unnormalised coordinates, clamp, 8-bit fractional weights). */
int gcb_setupTexture(gcb_ctx* ctx, int dx, int dy, int dz);
int gcb_copytotexture(gcb_ctx* ctx, float* d_phi, gcb_pitched_ptr data_ptr, int NX, int NY, int NZ);
int gcb_updateTexture(gcb_ctx* ctx, gcb_pitched_ptr data_ptr);
int gcb_deleteTexture(gcb_ctx* ctx);

/* Multitopo::unit_lattice, spectrum part (src/main.cu:3577-3706 with Fft_lattice::fft_func / fft_scalar / fft_fill,
 * src/lattice_files/Fft_lattice.cu:107-236): the (2*range_st+1)^3 lowest Fourier coefficients of the unit cell, divided by
 * the point count, in the order k, j, i = -range_st..range_st (i fastest) -- the reference's `lattice_data` (device
 * float2[(2*range_st+1)^3]), i.e. the c_h of the spatially varying lattice.  Evaluated directly (no FFT); agrees with the
 * reference's cuFFT route to ~1e-6 of the largest coefficient. */
int gcb_unit_lattice_spectrum(gcb_ctx* ctx, const float* d_unit_cell, int Nxu, int Nyu, int Nzu, int range_st, void* d_lattice_data);

/* ---- SVL phase solve.  Gratings::finding_phi (src/lattice_files/Gratings.h:40-41, Gratings.cu:100-417, :1015-1025)
 * and Gratings::GPUCG_lattice (Gratings.h:43, Gratings.cu:875-974; NX/NY/NZ are members there, arguments here) as
 * Multitopo::spatial_lattice_run calls them per harmonic (main.cu:3959-3962).  Same arithmetic, bit for bit (expression
 * contraction and reduction trees of the reference build); the CG scalars stay on the device, the host polls a counter.
 * FinalIter / FinalRes as the reference returns them. */
int gcb_finding_phi(gcb_ctx* ctx, float* d_phi, float* d_period, int x_dim, int y_dim, int z_dim, int i, int j, int k, float dx, float dy, float dz,
                    char latticetype_one, int unform_type, float const_peirod, float x_period, float y_period, float z_period, float lcon, float lcon_1,
                    int sinewave_zaxis);
int gcb_GPUCG_lattice(gcb_ctx* ctx, float* d_phi, int NX, int NY, int NZ, int iter, int OptIter, float EndRes, int* FinalIter, float* FinalRes);
/* fused: right-hand sides and CG of ALL harmonics at once.  d_phi_all: device float[nharm][z][y][x] (output), ijk: host int[3*nharm]
 * (the (i, j, k) of each harmonic), FinalIter / FinalRes: host arrays [nharm] (may be NULL).  Every harmonic performs exactly
 * the iterations the per-harmonic reference loop would, so each grid equals gcb_finding_phi + gcb_GPUCG_lattice bit for bit. */
int gcb_svl_phase_solve(gcb_ctx* ctx, float* d_phi_all, float* d_period, int nharm, const int* ijk, int x_dim, int y_dim, int z_dim, float dx, float dy, float dz,
                        char latticetype_one, int unform_type, float const_peirod, float x_period, float y_period, float z_period, float lcon, float lcon_1,
                        int sinewave_zaxis, int iter, float EndRes, int* FinalIter, float* FinalRes);

/* File_output::file_write_obj (src/File_output.h:38, File_output.cu:5-81): d_pos device float4[totalVerts].
 * Same file bytes as the reference writer; the weld, the face filter and the text formatting run on the GPU
 * (GCB_OPT_OBJ_HOST selects the single-thread host restatement). */
int gcb_file_write_obj(gcb_ctx* ctx, void* d_pos, unsigned int totalVerts, const char* filename);

/* ------------------------------------------------------------------ fused entry points (no reference twin) */
/* Each is validated against the composition of the legacy calls it replaces. */

/* Slab geometry for multi-GPU z-slab sharding.  A rank owns cell layers
 * [z0, z0 + nz_local - 1) of a global grid of gnz point layers and holds nz_local point layers
 * (its cells' +z halo plane included).  Single GPU: z0 = 0, gnz = nz_local. */
typedef struct gcb_slab { unsigned int z0, gnz; } gcb_slab;

/* SVL field: for h in [0,nh): svl += cos(phi_h)*re_h - sin(phi_h)*im_h, phi_h trilinear from the
 * control grid (replaces nh x {copytotexture, updateTexture, grating, svl}; main.cu:3949-3970).
 * d_phi: nh control grids of (cx,cy,cz_local) floats back to back; the control slab must cover the
 * coarse planes the fine slab samples: control plane index = floor((z0+z)*dz) - cz0 (and +1).
 * coef: nh float2 on the HOST.  If accumulate == 0 the field starts from 0 (cudaMemset in
 * check_lattice, main.cu:4058).  d_minmax (device float[2], optional) receives min/max of the
 * slab's field exactly as the reference's reduction defines them (seeded with 0). */
int gcb_svl_field(gcb_ctx* ctx, float* d_svl, const float* d_phi, int nh, const float* coef_host, int cx, int cy, int cz_local,
    int cz0, int NX2, int NY2, int NZ2_local, gcb_slab slab, float dx, float dy, float dz, int accumulate, float* d_minmax);

/* Same field from HOST control grids (pinned or pageable): the grids are uploaded in harmonic batches on a copy stream
 * while the kernel consumes the previous batch; bit-identical to gcb_svl_field.  d_phi_scratch: device float[nh*cx*cy*cz_local]. */
int gcb_svl_field_host(gcb_ctx* ctx, float* d_svl, const float* h_phi, float* d_phi_scratch, int nh, const float* coef_host, int cx, int cy,
    int cz_local, int cz0, int NX2, int NY2, int NZ2_local, gcb_slab slab, float dx, float dy, float dz, float* d_minmax);

/* min/max of a device array with the reference's semantics (result on host). */
int gcb_minmax(gcb_ctx* ctx, const float* d_in, size_t n, float* lo, float* hi);

/* Band-lattice extraction straight from the raw field: fuses device_bufferfour
 * (k = (f-a)/(b-a), domain faces forced to 0, band mask) into the marching-cubes kernel.
 * Equivalent to GPU_buffer_normalise_four + computeIsosurface_lattice(one) with vol_two = 0,
 * iso1 = iso2 = 0.  a, b: global min/max.  pos/norm: device float4[maxVerts].
 * d_compVoxelArray may be NULL.  Counts are 64-bit; vertex order is ascending global cell id.
 * count_only != 0: classify and count only (no mesh written). */
int gcb_extract_band_raw(gcb_ctx* ctx, const float* d_field, float a, float b, float isoValue, float isovalue1, float isovalue2,
    gcb_uint3 gridSizeLocal, gcb_slab slab, gcb_float3 voxelSize, gcb_float3 gridcenter, void* pos, void* norm,
    unsigned long long maxVerts, unsigned int* d_compVoxelArray, int count_only, unsigned long long* activeVoxels,
    unsigned long long* totalVerts);

/* Same, with the normalisation range read from DEVICE memory (d_minmax: {min, max}, e.g. the pair gcb_svl_field left there, or the
 * result of an NCCL all-reduce over ranks): no host round trip between the field and the extraction. */
int gcb_extract_band_raw_dev(gcb_ctx* ctx, const float* d_field, const float* d_minmax, float isoValue, float isovalue1, float isovalue2,
    gcb_uint3 gridSizeLocal, gcb_slab slab, gcb_float3 voxelSize, gcb_float3 gridcenter, void* pos, void* norm,
    unsigned long long maxVerts, unsigned int* d_compVoxelArray, int count_only, unsigned long long* activeVoxels,
    unsigned long long* totalVerts);

/* Whole config-3/4 pipeline on one rank, device-resident inputs: gcb_svl_field -> (a,b given by
 * caller or computed locally when use_local_minmax != 0) -> gcb_extract_band_raw. */
int gcb_svl_lattice(gcb_ctx* ctx, float* d_svl_scratch, const float* d_phi, int nh, const float* coef_host, int cx, int cy, int cz,
    int NX2, int NY2, int NZ2, float dx, float dy, float dz, float isoValue, float isovalue1, float isovalue2,
    gcb_float3 voxelSize, gcb_float3 gridcenter, void* pos, void* norm, unsigned long long maxVerts,
    unsigned long long* activeVoxels, unsigned long long* totalVerts, float* minmax_out);

/* Same pipeline with HOST inputs (pinned or pageable): copies the control grids host->device
 * into d_phi_scratch, runs gcb_svl_lattice, returns the counts.  This is the end-to-end call
 * bench.py times as `e2e`. */
int gcb_svl_lattice_host(gcb_ctx* ctx, const float* h_phi, float* d_phi_scratch, float* d_svl_scratch, int nh, const float* coef_host,
    int cx, int cy, int cz, int NX2, int NY2, int NZ2, float dx, float dy, float dz, float isoValue, float isovalue1,
    float isovalue2, gcb_float3 voxelSize, gcb_float3 gridcenter, void* pos, void* norm, unsigned long long maxVerts,
    unsigned long long* activeVoxels, unsigned long long* totalVerts, float* minmax_out);

/* Fused unit-lattice path (BASELINE config 1; Multitopo::display_unit_lattice, main.cu:4113-4132): the composition
 *   GPU_buffer_normalise_buffer(f, f) -> GPU_buffer_normalise_four(f, mask, k, isovalue1, isovalue2) ->
 *   computeIsosurface_latticeone(mask, ..., k, isovalue1, isovalue2)
 * on a RAW field in one call: one reduction of the field's true range, both normalisations applied inside the extraction kernel,
 * one synchronisation.  Counts and mesh are bit-identical to that composition.  ranges_out (host float[4], optional):
 * {a, b} of the first normalisation, {a2, b2} of the second.  d_compVoxelArray may be NULL. */
int gcb_band_lattice_from_raw(gcb_ctx* ctx, const float* d_raw_field, gcb_uint3 gridSize, float isoValue, float isovalue1, float isovalue2,
    gcb_float3 voxelSize, gcb_float3 gridcenter, void* pos, void* norm, unsigned long long maxVerts, unsigned int* d_compVoxelArray,
    unsigned long long* activeVoxels, unsigned long long* totalVerts, float* ranges_out);
/* ... with Fft_lattice::create_lattice in front: the unit cell of TPMS type 0..5 is written to d_field_scratch (device float[N]) and
 * its range reduced by the same kernel. */
int gcb_tpms_lattice(gcb_ctx* ctx, float* d_field_scratch, unsigned int lattice_type_index, gcb_uint3 gridSize, float isoValue, float isovalue1,
    float isovalue2, gcb_float3 voxelSize, gcb_float3 gridcenter, void* pos, void* norm, unsigned long long maxVerts,
    unsigned int* d_compVoxelArray, unsigned long long* activeVoxels, unsigned long long* totalVerts, float* ranges_out);

/* Fused density-surface path (BASELINE config 5; Multitopo::toprun, main.cu:3060-3109): Interpolations::copytotexture + updateTexture +
 * Gratings::refine + Isosurface::computeIsosurface_2 with an all-zero vol_topo / d_result, in one call and one synchronisation.
 * d_coarse: device float[cz][cy][cx]; d_density_fine: device float[NZ2*NY2*NX2], receives the upsampled density (as refine leaves it).
 * Counts and mesh equal that sequence bit for bit (norm.w = 0). */
int gcb_density_surface(gcb_ctx* ctx, const float* d_coarse, int cx, int cy, int cz, float* d_density_fine, int NX2, int NY2, int NZ2, float dx, float dy,
    float dz, float isoValue, gcb_float3 voxelSize, gcb_float3 gridcenter, void* pos, void* norm, unsigned long long maxVerts,
    unsigned int* d_compVoxelArray, unsigned long long* activeVoxels, unsigned long long* totalVerts);

/* Fused "add primitive" action (BASELINE config 2; Multitopo::show_model, main.cu:3304-3465): Modelling::<primitive>(d_field, ...) followed
 * by Isosurface::copy_parameter(..., vol_one, d_field, ..., obj_union, obj_diff, obj_intersect) (fixed / dynamic off) as ONE call.  kind:
 * gcb_primitive_kind; center / aux (Euler angles; the axis for GCB_PRIM_LINE) / flag (shell, disc) and params[] are the primitive's arguments
 * in the order of its own gcb_* entry point: sphere {radius, thickness}, line {radius, thickness_radial, thickness_axial}, cuboid {x, y, z
 * width}, cuboid_shell {x, y, z width, thickness}, torus {torus radius, circle radius}, cone {base radius, height}, cone_frustum {top radius,
 * bottom radius, height}, pyramid_frustum {x base, x top, y height, z base, z top}.  vol_one and (when given) d_field end up bit for bit
 * as after the two calls.  Sphere and the two cuboids on grids with Nx a multiple of four are evaluated INSIDE the retain kernel (one
 * pass, no field round trip; d_field may then be NULL = do not store the field); every other case runs the two kernels back to back
 * and needs d_field. */
enum gcb_primitive_kind {
    GCB_PRIM_SPHERE = 0, GCB_PRIM_LINE = 1, GCB_PRIM_CUBOID = 2, GCB_PRIM_CUBOID_SHELL = 3, GCB_PRIM_TORUS = 4, GCB_PRIM_CONE = 5,
    GCB_PRIM_CONE_FRUSTUM = 6, GCB_PRIM_PYRAMID_FRUSTUM = 7
};
int gcb_csg_retain_primitive(gcb_ctx* ctx, int kind, gcb_float3 center, gcb_float3 aux, const float* params, int nparams, int flag,
    float* d_field, gcb_grid_points* vol_one, int Nx, int Ny, int Nz, float dx, float dy, float dz, float isoValue, int obj_union, int obj_diff,
    int obj_intersect);

/* Two-deep job pipeline of gcb_svl_lattice_host: _submit only enqueues (H2D copies on the library's copy stream, field,
 * reduction, extraction and the read-back of the counts on the context's stream) and returns; _wait blocks until that slot's job
 * is complete and hands back its counts.  With jobs alternating between slot 0 and slot 1, the control grids of job i+1 cross
 * PCIe while job i computes.  Each slot needs its OWN d_phi_scratch (it is written while the other slot's job reads its own);
 * d_svl_scratch and pos / norm may be shared between the slots (their uses are ordered on the context's stream -- a shared mesh
 * buffer holds the mesh of the LAST submitted job).  h_phi must stay valid and unchanged until _wait returns.  Results equal
 * the blocking call's bit for bit. */
int gcb_svl_lattice_host_submit(gcb_ctx* ctx, int slot, const float* h_phi, float* d_phi_scratch, float* d_svl_scratch, int nh, const float* coef_host,
    int cx, int cy, int cz, int NX2, int NY2, int NZ2, float dx, float dy, float dz, float isoValue, float isovalue1, float isovalue2,
    gcb_float3 voxelSize, gcb_float3 gridcenter, void* pos, void* norm, unsigned long long maxVerts);
int gcb_svl_lattice_host_wait(gcb_ctx* ctx, int slot, unsigned long long* activeVoxels, unsigned long long* totalVerts, float* minmax_out);
/* The job pipeline for ONE z-slab of a sharded lattice (one process per GPU): the normalisation range is the range over ALL ranks, so a
 * job is enqueued in two halves.  _submit_field: H2D copies of the rank's control planes + field of the slab + its local {min, max} into
 * d_minmax (device float[2], the caller's).  The caller then reduces d_minmax over the ranks ON THE CONTEXT'S STREAM (e.g. an NCCL
 * all-reduce, which is stream-ordered: no host synchronisation) into d_ab and calls _submit_extract, which enqueues the extraction with
 * the range read from d_ab and the read-back of the counts.  gcb_svl_lattice_host_wait(slot) completes the job (minmax_out = the
 * contents of d_ab).  Arguments as gcb_svl_field_host / gcb_extract_band_raw_dev; the same buffer rules as above. */
int gcb_svl_slab_host_submit_field(gcb_ctx* ctx, int slot, const float* h_phi, float* d_phi_scratch, float* d_svl_scratch, int nh, const float* coef_host,
    int cx, int cy, int cz_local, int cz0, int NX2, int NY2, int NZ2_local, gcb_slab slab, float dx, float dy, float dz, float* d_minmax);
int gcb_svl_slab_host_submit_extract(gcb_ctx* ctx, int slot, const float* d_svl_scratch, const float* d_ab, float isoValue, float isovalue1,
    float isovalue2, gcb_uint3 gridSizeLocal, gcb_slab slab, gcb_float3 voxelSize, gcb_float3 gridcenter, void* pos, void* norm,
    unsigned long long maxVerts);

/* ------------------------------------------------------------------ several GPUs from one host process (z-slab sharding)
 * The reference is single-GPU; the scheme is BASELINE.json's: rank r owns cell layers [z0_r, z1_r) (gcb_slab_bounds, cuts aligned
 * to 2), holds / evaluates point layers z0_r .. z1_r, the global min/max joins the ranks between field and extraction, counts are
 * exclusive-scanned into global vertex offsets, and the rank meshes concatenated in rank order equal the single-GPU mesh byte for
 * byte.  gcb_multi_* owns one context + one stream per rank and enables peer access between the devices; `devices` may name the same
 * device more than once (slabs processed on one GPU, e.g. for tests).  Per-rank arguments are HOST arrays of n entries whose
 * elements are device pointers ON that rank's device.  All work is enqueued without host synchronisation; the calls return after
 * the counts of every rank have arrived. */
typedef struct gcb_multi gcb_multi;
int gcb_multi_create(gcb_multi** m, int n, const int* devices);
int gcb_multi_destroy(gcb_multi* m);
int gcb_multi_size(gcb_multi* m);
gcb_ctx* gcb_multi_ctx(gcb_multi* m, int rank);      /* the rank's context (its stream is a non-blocking stream of the rank's device) */
int gcb_multi_device(gcb_multi* m, int rank);
const char* gcb_multi_last_error(gcb_multi* m);
float gcb_multi_last_ms(gcb_multi* m);               /* device time of the last sharded call: max over ranks of (first launch .. counts) */
/* cell layers [z0, z1) of `rank` of a grid with gnz point layers; same cuts as gpucadforam_b200/sharding.py slab_bounds */
int gcb_slab_bounds(unsigned int gnz, int world, int rank, unsigned int align, unsigned int* z0, unsigned int* z1);
/* control planes [c0, c1] a fine slab with point layers z0 .. z1 samples at upsampling ratio `ratio` */
int gcb_control_slab(unsigned int z0, unsigned int z1, int ratio, int cz_global, int* c0, int* c1);

/* BASELINE configs 3 / 4 on n ranks: gcb_svl_lattice sharded.  d_phi[r]: rank r's control slab (planes cz0[r] .. cz0[r] + cz_local[r] - 1
 * of all harmonics), d_svl[r]: field scratch of NX2 * NY2 * (z1_r - z0_r + 1) floats.  The min/max exchange is a one-warp kernel per
 * rank reading every rank's pair through peer-mapped pointers.  count_only != 0: counts only (pos / norm / max_verts may be NULL). */
int gcb_multi_svl_lattice(gcb_multi* m, float* const* d_svl, const float* const* d_phi, int nh, const float* coef_host, int cx, int cy,
    const int* cz_local, const int* cz0, int NX2, int NY2, unsigned int gnz, float dx, float dy, float dz, float isoValue, float isovalue1,
    float isovalue2, gcb_float3 voxelSize, gcb_float3 gridcenter, void* const* pos, void* const* norm, const unsigned long long* max_verts,
    int count_only, unsigned long long* active, unsigned long long* verts, unsigned long long* vert_offsets, float* minmax_out);

/* Isosurface::computeIsosurface_2 on STORED fields sharded over n ranks (BASELINE config 5 on several GPUs).  Rank r's arrays hold its
 * OWNED point layers only -- z0_r .. z1_r - 1, the last rank also its final layer; the +z halo layer is not copied: the extraction
 * kernel stages it from the upper neighbour's arrays where they lie (peer memory over NVLink).  vol_topo / d_result may be NULL
 * (treated as zeros).  Vertex z and d_compVoxelArray ids are global. */
int gcb_multi_computeIsosurface_2(gcb_multi* m, gcb_grid_points* const* vol_topo, float* const* vol_two, float* const* d_result,
    gcb_uint3 gridSizeGlobal, gcb_float3 voxelSize, gcb_float3 gridcenter, float isoValue, float isovalue1, void* const* pos, void* const* norm,
    const unsigned long long* max_verts, unsigned int* const* d_compVoxelArray, unsigned long long* active, unsigned long long* verts,
    unsigned long long* vert_offsets, unsigned long long* active_offsets);

#ifdef __cplusplus
}
#endif
#endif /* GPUCAD_B200_H */
