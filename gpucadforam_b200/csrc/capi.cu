// capi.cu -- extern "C" boundary of libgpucad_b200 (declared in include/gpucad_b200.h).
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include <cstdio>
#include <cstring>
#include <new>

namespace gcb {

int fail(Ctx* c, const char* what, cudaError_t e) {
    if (c) c->err = std::string(what) + ": " + cudaGetErrorString(e);
    return 1;
}
int fail_msg(Ctx* c, const std::string& msg) {
    if (c) c->err = msg;
    return 1;
}


// Legacy extraction: optional byte-granular memset of Isosurface.cu:120-121, then the fused kernel.
static int run_legacy(Ctx* c, McArgs& a, void* pos, void* norm, unsigned maxVerts, unsigned* d_verts, unsigned* d_vertsScan, unsigned* d_occ,
                      unsigned* d_occScan, unsigned* d_comp, unsigned* activeVoxels, unsigned* totalVerts, bool memset_out) {
    a.pos = (float4*)pos;
    a.norm = (float4*)norm;
    a.max_verts = maxVerts;
    a.comp = d_comp;
    a.gz0 = 0;
    a.gnz = a.nz;
    a.count_only = 0;
    const bool fill = (c->options & GCB_OPT_FILL_STAGE_ARRAYS) && d_verts && d_vertsScan && d_occ && d_occScan;
    a.st_verts = fill ? d_verts : nullptr;
    a.st_occ = fill ? d_occ : nullptr;
    a.st_verts_scan = fill ? d_vertsScan : nullptr;
    a.st_occ_scan = fill ? d_occScan : nullptr;
    if (memset_out && (c->options & GCB_OPT_LEGACY_MEMSET)) {
        // The reference clears maxVerts BYTES (not vertices) of both buffers after it knows activeVoxels > 0.
        // Clearing before the kernel is equivalent: every byte the kernel does not overwrite is identical.
        // (When activeVoxels == 0 the reference leaves the buffers untouched; consumers honour totalVerts.)
        GCB_CHECK(c, cudaMemsetAsync(pos, 0, maxVerts, c->stream));
        GCB_CHECK(c, cudaMemsetAsync(norm, 0, maxVerts, c->stream));
    }
    unsigned long long act = 0, verts = 0;
    if (int r = launch_extract(c, a, &act, &verts)) return r;
    if (activeVoxels) *activeVoxels = (unsigned)act;
    if (totalVerts) *totalVerts = (unsigned)verts;
    return 0;
}


} // namespace gcb

using namespace gcb;

#define CTX(c) Ctx* C = reinterpret_cast<Ctx*>(c); if (!C) return 1

extern "C" {

int gcb_create(gcb_ctx** out, int device, void* stream) {
    if (!out) return 1;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device >= n) return 2;  // no CPU fallback
    if (cudaSetDevice(device) != cudaSuccess) return 2;
    Ctx* c = new (std::nothrow) Ctx();
    if (!c) return 1;
    c->device = device;
    c->stream = (cudaStream_t)stream;
    cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device);
    bool ok = cudaMalloc(&c->d_tile_counter, 16) == cudaSuccess && cudaMalloc(&c->d_totals, 16) == cudaSuccess &&
              cudaMallocHost(&c->h_totals, 16) == cudaSuccess && cudaMalloc(&c->d_minmax, 16) == cudaSuccess &&
              cudaMallocHost(&c->h_minmax, 16) == cudaSuccess;
    for (int i = 0; i < 4 && ok; ++i) ok = cudaEventCreate(&c->ev[i]) == cudaSuccess;
    if (!ok) { gcb_destroy(reinterpret_cast<gcb_ctx*>(c)); return 3; }  // releases whatever was allocated before the failure
    *out = reinterpret_cast<gcb_ctx*>(c);
    return 0;
}
int gcb_destroy(gcb_ctx* ctx) {
    CTX(ctx);
    int prev = -1;
    cudaGetDevice(&prev);
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev};  // leave the caller's current device as it was
    cudaSetDevice(C->device);
    cudaFree(C->d_status); cudaFree(C->d_tile_counter); cudaFree(C->d_totals); cudaFreeHost(C->h_totals);
    cudaFree(C->d_minmax); cudaFreeHost(C->h_minmax); cudaFree(C->d_tex); cudaFree(C->d_coef); cudaFree(C->d_range4); cudaFree(C->d_tab);
    cudaFree(C->d_tri); cudaFree(C->d_nverts);
    if (C->copy_stream) { cudaStreamDestroy(C->copy_stream); for (int i = 0; i < Ctx::kBatches; ++i) cudaEventDestroy(C->copy_ev[i]); }
    if (C->aux_stream) { cudaStreamDestroy(C->aux_stream); cudaEventDestroy(C->aux_ev[0]); cudaEventDestroy(C->aux_ev[1]); }
    for (int i = 0; i < 4; ++i) if (C->ev[i]) cudaEventDestroy(C->ev[i]);
    for (auto& sl : C->slot) {
        if (sl.h_totals) cudaFreeHost(sl.h_totals);
        if (sl.h_minmax) cudaFreeHost(sl.h_minmax);
        if (sl.d_minmax) cudaFree(sl.d_minmax);
        if (sl.field_done) cudaEventDestroy(sl.field_done);
        if (sl.job_done) cudaEventDestroy(sl.job_done);
    }
    delete C;
    return 0;
}
const char* gcb_last_error(gcb_ctx* ctx) { Ctx* C = reinterpret_cast<Ctx*>(ctx); return C ? C->err.c_str() : "null context"; }
int gcb_set_stream(gcb_ctx* ctx, void* stream) { CTX(ctx); C->stream = (cudaStream_t)stream; return 0; }
int gcb_set_options(gcb_ctx* ctx, unsigned int flags) { CTX(ctx); C->options = flags; return 0; }
unsigned long long gcb_launch_count(gcb_ctx* ctx) { Ctx* C = reinterpret_cast<Ctx*>(ctx); return C ? C->launches : 0; }
void gcb_reset_launch_count(gcb_ctx* ctx) { Ctx* C = reinterpret_cast<Ctx*>(ctx); if (C) C->launches = 0; }
int gcb_enable_kernel_timing(gcb_ctx* ctx, int on) { CTX(ctx); C->timing = on != 0; return 0; }
float gcb_last_extract_kernel_ms(gcb_ctx* ctx) {
    Ctx* C = reinterpret_cast<Ctx*>(ctx);
    if (!C || !C->extract_timed) return -1.f;
    float ms = -1.f;
    if (cudaEventSynchronize(C->ev[1]) != cudaSuccess || cudaEventElapsedTime(&ms, C->ev[0], C->ev[1]) != cudaSuccess) return -1.f;
    return ms;
}
float gcb_last_field_kernel_ms(gcb_ctx* ctx) {
    Ctx* C = reinterpret_cast<Ctx*>(ctx);
    if (!C || !C->field_timed) return -1.f;
    float ms = -1.f;
    if (cudaEventSynchronize(C->ev[3]) != cudaSuccess || cudaEventElapsedTime(&ms, C->ev[2], C->ev[3]) != cudaSuccess) return -1.f;
    return ms;
}

void gcb_tables(unsigned int* tri, unsigned int* nverts) { host_tables(tri, nverts); }

int gcb_allocateTextures_s(gcb_ctx* ctx, unsigned int** d_triTable, unsigned int** d_numVertsTable) {
    CTX(ctx);
    unsigned tri[256 * 16], nv[256];
    host_tables(tri, nv);
    if (!C->d_tri) {
        GCB_CHECK(C, cudaMalloc(&C->d_tri, sizeof tri));
        GCB_CHECK(C, cudaMalloc(&C->d_nverts, sizeof nv));
    }
    GCB_CHECK(C, cudaMemcpy(C->d_tri, tri, sizeof tri, cudaMemcpyHostToDevice));
    GCB_CHECK(C, cudaMemcpy(C->d_nverts, nv, sizeof nv, cudaMemcpyHostToDevice));
    if (d_triTable) *d_triTable = C->d_tri;
    if (d_numVertsTable) *d_numVertsTable = C->d_nverts;
    return 0;
}
int gcb_destroyAllTextureObjects(gcb_ctx* ctx) { CTX(ctx); return 0; }

// ------------------------------------------------------------------ extraction, legacy signatures
int gcb_computeIsosurface(gcb_ctx* ctx, float* vol, gcb_uint3 raster_grid, void* pos, void* norm, float isoValue, unsigned int numVoxels,
                          unsigned int* d_voxelVerts, unsigned int* d_voxelVertsScan, unsigned int* d_voxelOccupied, unsigned int* d_voxelOccupiedScan,
                          gcb_uint3 gridSize, gcb_uint3 gridSizeShift, gcb_uint3 gridSizeMask, gcb_float3 voxelSize, gcb_float3 gridcenter,
                          unsigned int* activeVoxels, unsigned int* totalVerts, unsigned int* d_compVoxelArray, unsigned int maxVerts,
                          gcb_grid_points* primitive_fixed, float* primitive_dynamic, float* topo_field, float* lattice_field, float iso1, float iso2,
                          int obj_union, int obj_diff, int obj_intersect, int primitive, int topo, int compute_lattice, int fixed, int dynamic,
                          int make_region, size_t* nfacets) {
    CTX(ctx);
    // dead parameters of the reference kernel: vol, raster_grid, topo_field, primitive, topo, compute_lattice
    (void)vol; (void)raster_grid; (void)numVoxels; (void)gridSizeShift; (void)gridSizeMask; (void)topo_field; (void)primitive; (void)topo; (void)compute_lattice;
    McArgs a;
    base_args(a, M_CSG, gridSize, voxelSize, gridcenter, isoValue);
    a.iso1 = iso1; a.iso2 = iso2;
    a.flags = (obj_union ? F_UNION : 0) | (obj_diff ? F_DIFF : 0) | (obj_intersect ? F_INTERSECT : 0) | (fixed ? F_FIXED : 0) | (dynamic ? F_DYNAMIC : 0) |
              (make_region ? F_MAKE_REGION : 0);
    a.f0 = primitive_dynamic;
    a.f1 = lattice_field;
    a.gp = (const GridPoint*)primitive_fixed;
    unsigned tv = 0;
    int r = run_legacy(C, a, pos, norm, maxVerts, d_voxelVerts, d_voxelVertsScan, d_voxelOccupied, d_voxelOccupiedScan, d_compVoxelArray, activeVoxels, &tv, true);
    if (r) return r;
    if (totalVerts) *totalVerts = tv;
    if (nfacets && tv) *nfacets = tv / 3;  // Isosurface.cu:115-116 (not written on the early-out path)
    return 0;
}

int gcb_computeIsosurface_region(gcb_ctx* ctx, void* pos, void* norm, float isoValue, unsigned int numVoxels, unsigned int* d_voxelVerts,
                                 unsigned int* d_voxelVertsScan, unsigned int* d_voxelOccupied, unsigned int* d_voxelOccupiedScan, gcb_uint3 gridSize,
                                 gcb_uint3 gridSizeShift, gcb_uint3 gridSizeMask, gcb_float3 voxelSize, gcb_float3 gridcenter, unsigned int* activeVoxels,
                                 unsigned int* totalVerts, unsigned int* d_compVoxelArray, unsigned int maxVerts, gcb_grid_points* vol_topo,
                                 gcb_grid_points* primitive_fixed, float* primitive_dynamic, float* topo_field, float* lattice_field, float iso1, float iso2,
                                 int obj_union, int obj_diff, int obj_intersect, int primitive, int topo, int compute_lattice, int fixed, int dynamic,
                                 int make_region, int show_region, int show_domain, gcb_triangle_metadata* triangle_data) {
    CTX(ctx);
    // dead parameters of the reference kernels (MarchingCubes_kernel.cu:1163-1167, :2222-2228): everything but the three flags below
    (void)numVoxels; (void)gridSizeShift; (void)gridSizeMask; (void)topo_field; (void)lattice_field; (void)iso1; (void)iso2; (void)obj_union; (void)obj_diff;
    (void)obj_intersect; (void)primitive; (void)topo; (void)compute_lattice; (void)fixed; (void)dynamic;
    // with none of the three flags the reference classifies an uninitialised cube index (:1214-1285): refuse instead of guessing
    if (!make_region && !show_region && !show_domain) return fail_msg(C, "computeIsosurface_region: one of make_region / show_region / show_domain must be set");
    if (!vol_topo || !primitive_fixed || !primitive_dynamic) return fail_msg(C, "computeIsosurface_region: null field");
    if (show_region && !triangle_data) return fail_msg(C, "computeIsosurface_region: show_region needs triangle_data");
    McArgs a;
    base_args(a, M_REGION, gridSize, voxelSize, gridcenter, isoValue);
    // precedence as the reference's if / else-if chain: show_region, then show_domain, then make_region
    a.flags = show_region ? F_SHOW_REGION : show_domain ? F_SHOW_DOMAIN : F_MAKE_REGION;
    a.f0 = primitive_dynamic;
    a.gp = (const GridPoint*)primitive_fixed;
    a.gp2 = (const GridPoint*)vol_topo;
    a.meta = (TriangleMetadata*)triangle_data;
    return run_legacy(C, a, pos, norm, maxVerts, d_voxelVerts, d_voxelVertsScan, d_voxelOccupied, d_voxelOccupiedScan, d_compVoxelArray, activeVoxels, totalVerts,
                      true);  // Isosurface.cu:217-218 clears maxVerts bytes like the CSG variant
}

static int lattice_common(Ctx* C, int mode, float* vol, void* pos, void* norm, float isoValue, unsigned* d_voxelVerts, unsigned* d_voxelVertsScan,
                          unsigned* d_voxelOccupied, unsigned* d_voxelOccupiedScan, gcb_uint3 gridSize, gcb_float3 voxelSize, gcb_float3 gridcenter,
                          unsigned* activeVoxels, unsigned* totalVerts, unsigned* d_compVoxelArray, unsigned maxVerts, float* vol_one, float* vol_two,
                          float isovalue1, float isovalue2, float iso1, float iso2) {
    McArgs a;
    base_args(a, mode, gridSize, voxelSize, gridcenter, isoValue);
    a.iso1 = isovalue1; a.iso2 = isovalue2; a.iso1b = iso1; a.iso2b = iso2;
    a.f0 = vol_one;  // k: interpolation field, TMA-staged
    a.f1 = vol;      // mask: classification field
    a.f2 = vol_two;
    // the lattice variants have no memset in the reference (Isosurface.cu:401-572)
    return run_legacy(C, a, pos, norm, maxVerts, d_voxelVerts, d_voxelVertsScan, d_voxelOccupied, d_voxelOccupiedScan, d_compVoxelArray, activeVoxels, totalVerts,
                      false);
}

int gcb_computeIsosurface_lattice(gcb_ctx* ctx, float* vol, void* pos, void* norm, float isoValue, unsigned int numVoxels, unsigned int* d_voxelVerts,
                                  unsigned int* d_voxelVertsScan, unsigned int* d_voxelOccupied, unsigned int* d_voxelOccupiedScan, gcb_uint3 gridSize,
                                  gcb_uint3 gridSizeShift, gcb_uint3 gridSizeMask, gcb_float3 voxelSize, gcb_float3 gridcenter, unsigned int* activeVoxels,
                                  unsigned int* totalVerts, unsigned int* d_compVoxelArray, unsigned int maxVerts, float* vol_one, float* vol_two,
                                  float isovalue1, float isovalue2, float iso1, float iso2) {
    CTX(ctx);
    (void)numVoxels; (void)gridSizeShift; (void)gridSizeMask;
    if (!vol || !vol_one || !vol_two) return fail_msg(C, "computeIsosurface_lattice: null field");
    return lattice_common(C, M_LATTICE, vol, pos, norm, isoValue, d_voxelVerts, d_voxelVertsScan, d_voxelOccupied, d_voxelOccupiedScan, gridSize, voxelSize,
                          gridcenter, activeVoxels, totalVerts, d_compVoxelArray, maxVerts, vol_one, vol_two, isovalue1, isovalue2, iso1, iso2);
}
int gcb_computeIsosurface_latticeone(gcb_ctx* ctx, float* vol, void* pos, void* norm, float isoValue, unsigned int numVoxels, unsigned int* d_voxelVerts,
                                     unsigned int* d_voxelVertsScan, unsigned int* d_voxelOccupied, unsigned int* d_voxelOccupiedScan, gcb_uint3 gridSize,
                                     gcb_uint3 gridSizeShift, gcb_uint3 gridSizeMask, gcb_float3 voxelSize, gcb_float3 gridcenter,
                                     unsigned int* activeVoxels, unsigned int* totalVerts, unsigned int* d_compVoxelArray, unsigned int maxVerts,
                                     float* vol_one, float isovalue1, float isovalue2) {
    CTX(ctx);
    (void)numVoxels; (void)gridSizeShift; (void)gridSizeMask;
    if (!vol || !vol_one) return fail_msg(C, "computeIsosurface_latticeone: null field");
    return lattice_common(C, M_LATTICE_ONE, vol, pos, norm, isoValue, d_voxelVerts, d_voxelVertsScan, d_voxelOccupied, d_voxelOccupiedScan, gridSize, voxelSize,
                          gridcenter, activeVoxels, totalVerts, d_compVoxelArray, maxVerts, vol_one, nullptr, isovalue1, isovalue2, 0.f, 0.f);
}

static int topo_common(Ctx* C, void* pos, void* norm, float isoValue, unsigned* d_voxelVerts, unsigned* d_voxelVertsScan, unsigned* d_voxelOccupied,
                       unsigned* d_voxelOccupiedScan, gcb_uint3 gridSize, gcb_float3 voxelSize, gcb_float3 gridcenter, unsigned* activeVoxels,
                       unsigned* totalVerts, unsigned* d_compVoxelArray, unsigned maxVerts, gcb_grid_points* vol_topo, float* vol_two, float isovalue1,
                       float* d_result, int disp, void* disp_two) {
    if (!vol_two) return fail_msg(C, "computeIsosurface_topo/_2: null density field");
    McArgs a;
    base_args(a, M_TOPO, gridSize, voxelSize, gridcenter, isoValue);
    a.iso1 = isovalue1;
    a.f0 = vol_two;
    a.f1 = d_result;
    a.gp = (const GridPoint*)vol_topo;
    a.disp = (const float4*)disp_two;
    a.flags = (disp && disp_two) ? F_DISP : 0;
    return run_legacy(C, a, pos, norm, maxVerts, d_voxelVerts, d_voxelVertsScan, d_voxelOccupied, d_voxelOccupiedScan, d_compVoxelArray, activeVoxels, totalVerts,
                      false);
}
int gcb_computeIsosurface_2(gcb_ctx* ctx, void* pos, void* norm, float isoValue, unsigned int numVoxels, unsigned int* d_voxelVerts,
                            unsigned int* d_voxelVertsScan, unsigned int* d_voxelOccupied, unsigned int* d_voxelOccupiedScan, gcb_uint3 gridSize,
                            gcb_uint3 gridSizeShift, gcb_uint3 gridSizeMask, gcb_float3 voxelSize, gcb_float3 gridcenter, unsigned int* activeVoxels,
                            unsigned int* totalVerts, unsigned int* d_compVoxelArray, unsigned int maxVerts, gcb_grid_points* vol_topo,
                            gcb_grid_points* vol_one, float* vol_two, float* d_solid, float isovalue1, float* d_result, void* triangle_data) {
    CTX(ctx);
    (void)numVoxels; (void)gridSizeShift; (void)gridSizeMask; (void)vol_one; (void)d_solid; (void)triangle_data;  // dead in the reference kernels (A-15)
    return topo_common(C, pos, norm, isoValue, d_voxelVerts, d_voxelVertsScan, d_voxelOccupied, d_voxelOccupiedScan, gridSize, voxelSize, gridcenter,
                       activeVoxels, totalVerts, d_compVoxelArray, maxVerts, vol_topo, vol_two, isovalue1, d_result, 0, nullptr);
}
int gcb_computeIsosurface_topo(gcb_ctx* ctx, void* pos, void* norm, float isoValue, unsigned int numVoxels, unsigned int* d_voxelVerts,
                               unsigned int* d_voxelVertsScan, unsigned int* d_voxelOccupied, unsigned int* d_voxelOccupiedScan, gcb_uint3 gridSize,
                               gcb_uint3 gridSizeShift, gcb_uint3 gridSizeMask, gcb_float3 voxelSize, gcb_float3 gridcenter, unsigned int* activeVoxels,
                               unsigned int* totalVerts, unsigned int* d_compVoxelArray, unsigned int maxVerts, gcb_grid_points* vol_topo,
                               gcb_grid_points* vol_one, float* vol_two, float* d_solid, float isovalue1, float* d_result, void* triangle_data, int disp,
                               void* disp_two) {
    CTX(ctx);
    (void)numVoxels; (void)gridSizeShift; (void)gridSizeMask; (void)vol_one; (void)d_solid; (void)triangle_data;
    return topo_common(C, pos, norm, isoValue, d_voxelVerts, d_voxelVertsScan, d_voxelOccupied, d_voxelOccupiedScan, gridSize, voxelSize, gridcenter,
                       activeVoxels, totalVerts, d_compVoxelArray, maxVerts, vol_topo, vol_two, isovalue1, d_result, disp, disp_two);
}

// End of a legacy call that returns nothing through host pointers: the reference wrappers end in cudaDeviceSynchronize
// (e.g. MarchingCubes_kernel.cu:458, :1074); with GCB_OPT_ASYNC_FIELDS the call only enqueues on the context's stream (the next
// call that reports counts or writes a file synchronises, and launch errors surface here through cudaGetLastError).
static int field_call_end(Ctx* C) {
    if (C->options & GCB_OPT_ASYNC_FIELDS) {
        GCB_CHECK(C, cudaGetLastError());
        return 0;
    }
    GCB_CHECK(C, cudaStreamSynchronize(C->stream));
    return 0;
}

int gcb_copy_parameter(gcb_ctx* ctx, unsigned int* voxel_verts, float isoValue, gcb_uint3 gridSize, gcb_uint3 gridSizeShift, gcb_uint3 gridSizeMask,
                       gcb_float3 voxelSize, unsigned int numVoxels, gcb_grid_points* vol_one, float* vol_two, float* vol_lattice, int fixed, int dynamic,
                       float iso1, float iso2, int obj_union, int obj_diff, int obj_intersect) {
    CTX(ctx);
    (void)voxel_verts; (void)gridSizeShift; (void)gridSizeMask; (void)voxelSize; (void)numVoxels; (void)fixed;
    int r = k_copy_parameter(C, (GridPoint*)vol_one, vol_two, vol_lattice, dynamic != 0, iso1, iso2, gridSize.x, gridSize.y, gridSize.z, isoValue,
                             obj_union != 0, obj_diff != 0, obj_intersect != 0);
    if (r) return r;
    return field_call_end(C);  // the reference wrapper ends in cudaDeviceSynchronize (:458)
}
int gcb_csg_retain_primitive(gcb_ctx* ctx, int kind, gcb_float3 center, gcb_float3 aux, const float* params, int nparams, int flag, float* d_field,
                             gcb_grid_points* vol_one, int Nx, int Ny, int Nz, float dx, float dy, float dz, float isoValue, int obj_union, int obj_diff,
                             int obj_intersect) {
    CTX(ctx);
    int r = k_csg_retain_primitive(C, kind, f3(center), f3(aux), params, nparams, flag, d_field, (GridPoint*)vol_one, Nx, Ny, Nz, dx, dy, dz, isoValue,
                                   obj_union != 0, obj_diff != 0, obj_intersect != 0);
    if (r) return r;
    return field_call_end(C);
}
int gcb_patch_topo_field(gcb_ctx* ctx, float* d_vec1, int Nx, int Ny, int Nz, gcb_grid_points* vol_one) {
    CTX(ctx);
    if (int r = k_patch_topo_field(C, d_vec1, Nx, Ny, Nz, (const GridPoint*)vol_one)) return r;
    return field_call_end(C);
}

// ------------------------------------------------------------------ fields, legacy signatures
#define SYNC_RET(expr)                  \
    do {                                \
        if (int r_ = (expr)) return r_; \
        return field_call_end(C);       \
    } while (0)

int gcb_distance_from_line(gcb_ctx* ctx, float* d, gcb_float3 center, gcb_float3 axis, float radius_1, float thickness_radial, float thickness_axial, int Nx,
                           int Ny, int Nz, float dx, float dy, float dz, int disc) {
    CTX(ctx);
    SYNC_RET(k_line(C, d, f3(center), f3(axis), radius_1, thickness_radial, thickness_axial, Nx, Ny, Nz, dx, dy, dz, disc != 0));
}
int gcb_sphere_with_center(gcb_ctx* ctx, float* d, gcb_float3 center, float radius_1, float thickness_wall, int Nx, int Ny, int Nz, float dx, float dy, float dz,
                           int shell) {
    CTX(ctx);
    SYNC_RET(k_sphere(C, d, f3(center), radius_1, thickness_wall, Nx, Ny, Nz, dx, dy, dz, shell != 0));
}
int gcb_cuboid(gcb_ctx* ctx, float* d, gcb_float3 center, gcb_float3 angles, float xw, float yw, float zw, int Nx, int Ny, int Nz, float dx, float dy, float dz) {
    CTX(ctx);
    SYNC_RET(k_cuboid(C, d, f3(center), f3(angles), xw, yw, zw, Nx, Ny, Nz, dx, dy, dz));
}
int gcb_cuboid_shell(gcb_ctx* ctx, float* d, gcb_float3 center, gcb_float3 angles, float xw, float yw, float zw, float thickness, int Nx, int Ny, int Nz,
                     float dx, float dy, float dz) {
    CTX(ctx);
    SYNC_RET(k_cuboid_shell(C, d, f3(center), f3(angles), xw, yw, zw, thickness, Nx, Ny, Nz, dx, dy, dz));
}
int gcb_torus_with_center(gcb_ctx* ctx, float* d, gcb_float3 center, gcb_float3 angles, float torus_radius, float torus_circle_radius, int Nx, int Ny, int Nz,
                          float dx, float dy, float dz) {
    CTX(ctx);
    SYNC_RET(k_torus(C, d, f3(center), f3(angles), torus_radius, torus_circle_radius, Nx, Ny, Nz, dx, dy, dz));
}
int gcb_cone_with_base_radius_height(gcb_ctx* ctx, float* d, gcb_float3 center, gcb_float3 angles, float base_radius, float cone_height, int Nx, int Ny, int Nz,
                                     float dx, float dy, float dz) {
    CTX(ctx);
    SYNC_RET(k_cone(C, d, f3(center), f3(angles), base_radius, cone_height, Nx, Ny, Nz, dx, dy, dz));
}
int gcb_cone_frustum(gcb_ctx* ctx, float* d, gcb_float3 center, gcb_float3 angles, float top_radius, float bottom_radius, float h, int Nx, int Ny, int Nz,
                     float dx, float dy, float dz) {
    CTX(ctx);
    SYNC_RET(k_cone_frustum(C, d, f3(center), f3(angles), top_radius, bottom_radius, h, Nx, Ny, Nz, dx, dy, dz));
}
int gcb_pyramid_frustum(gcb_ctx* ctx, float* d, gcb_float3 center, gcb_float3 angles, float xb, float xt, float yh, float zb, float zt, int Nx, int Ny, int Nz,
                        float dx, float dy, float dz) {
    CTX(ctx);
    SYNC_RET(k_pyramid_frustum(C, d, f3(center), f3(angles), xb, xt, yh, zb, zt, Nx, Ny, Nz, dx, dy, dz));
}
int gcb_create_lattice(gcb_ctx* ctx, float* d, unsigned int NX, unsigned int NY, unsigned int NZ, unsigned int size, unsigned int type) {
    CTX(ctx);
    (void)size;
    SYNC_RET(k_create_lattice(C, d, NX, NY, NZ, type));
}
int gcb_unit_lattice_spectrum(gcb_ctx* ctx, const float* d_unit_cell, int Nxu, int Nyu, int Nzu, int range_st, void* d_lattice_data) {
    CTX(ctx);
    SYNC_RET(k_unit_spectrum(C, d_unit_cell, Nxu, Nyu, Nzu, range_st, (float2*)d_lattice_data));
}
int gcb_GPU_buffer_normalise_buffer(gcb_ctx* ctx, float* d_vec1, float* d_vec2, int n) {
    CTX(ctx);
    // min/max stay in device memory between the reduction and the normalisation (one sync per call instead of two)
    if (int r = k_minmax_device(C, d_vec1, (size_t)n)) return r;
    SYNC_RET(k_normalise(C, d_vec1, d_vec2, (size_t)n, 0.f, 0.f, C->d_minmax));
}
int gcb_GPU_buffer_normalise_four(gcb_ctx* ctx, float* dataone, float* datatwo, float* datathree, size_t size, int Nx, int Ny, int Nz, float isoval_1,
                                  float isoval_2) {
    CTX(ctx);
    if (int r = k_minmax_device(C, dataone, size)) return r;
    SYNC_RET(k_normalise_four(C, dataone, datatwo, datathree, Nx, Ny, Nz, 0.f, 0.f, isoval_1, isoval_2, C->d_minmax));
}
int gcb_GPU_buffer_normalise_three(gcb_ctx* ctx, float* dataone, float* datatwo, size_t size, float a1, float b1) {
    CTX(ctx);
    if (int r = k_minmax_device(C, dataone, size)) return r;
    SYNC_RET(k_normalise_three(C, dataone, datatwo, size, a1, b1, C->d_minmax));
}
int gcb_period_data(gcb_ctx* ctx, float* d_period, int NX, int NY, int NZ, float dx, float dy, float dz, float mean_x, float mean_y, float mean_z, char axis) {
    CTX(ctx);
    SYNC_RET(k_period_angle(C, d_period, NX, NY, NZ, dx, dy, dz, mean_x, mean_y, mean_z, (int)axis, false));
}
int gcb_angle_data(gcb_ctx* ctx, float* d_theta, int NX, int NY, int NZ, float dx, float dy, float dz, float mean_x, float mean_y, float mean_z, char axis) {
    CTX(ctx);
    SYNC_RET(k_period_angle(C, d_theta, NX, NY, NZ, dx, dy, dz, mean_x, mean_y, mean_z, (int)axis, true));
}
int gcb_minmax(gcb_ctx* ctx, const float* d_in, size_t n, float* lo, float* hi) { CTX(ctx); return k_minmax(C, d_in, n, lo, hi); }

int gcb_grating(gcb_ctx* ctx, void* dvol, int NX2, int NY2, int NZ2, float dx2, float dy2, float dz2) {
    CTX(ctx);
    if (!C->d_tex) return fail_msg(C, "grating: setupTexture/updateTexture not called");
    SYNC_RET(k_grating(C, C->d_tex, C->tex_x, C->tex_y, C->tex_z, (float2*)dvol, NX2, NY2, NZ2, dx2, dy2, dz2));
}
int gcb_refine(gcb_ctx* ctx, float* dvol, int NX2, int NY2, int NZ2, float dx, float dy, float dz) {
    CTX(ctx);
    if (!C->d_tex) return fail_msg(C, "refine: setupTexture/updateTexture not called");
    SYNC_RET(k_refine(C, C->d_tex, C->tex_x, C->tex_y, C->tex_z, dvol, NX2, NY2, NZ2, dx, dy, dz));
}
int gcb_svl(gcb_ctx* ctx, float* d_svl, void* d_grating, int NX, int NY, int NZ, int indxx, void* data_fft) {
    CTX(ctx);
    SYNC_RET(k_svl(C, d_svl, (const float2*)d_grating, (size_t)NX * NY * NZ, indxx, (const float2*)data_fft));
}
int gcb_topo_field(gcb_ctx* ctx, float* topo_field, float* isosurf, float volfrac, int NX, int NY, int NZ) {
    CTX(ctx);
    SYNC_RET(k_topo_field(C, topo_field, isosurf, volfrac, (size_t)NX * NY * NZ));
}
int gcb_primitive_field(gcb_ctx* ctx, gcb_grid_points* primitive_field, float* primitive_active, float* isosurf, float isoval, int fixed, int active, int NX,
                        int NY, int NZ) {
    CTX(ctx);
    (void)isoval;
    SYNC_RET(k_primitive_field(C, (const GridPoint*)primitive_field, primitive_active, isosurf, (size_t)NX * NY * NZ, fixed != 0, active != 0));
}

int gcb_setupTexture(gcb_ctx* ctx, int dx, int dy, int dz) {
    CTX(ctx);
    if (C->d_tex) { cudaFree(C->d_tex); C->d_tex = nullptr; }
    C->tex_x = dx; C->tex_y = dy; C->tex_z = dz;
    GCB_CHECK(C, cudaMalloc(&C->d_tex, (size_t)dx * dy * dz * sizeof(float)));
    return 0;
}
int gcb_copytotexture(gcb_ctx* ctx, float* d_phi, gcb_pitched_ptr data_ptr, int NX, int NY, int NZ) {
    CTX(ctx);
    SYNC_RET(k_copy_to_pitched(C, d_phi, data_ptr, NX, NY, NZ));
}
int gcb_updateTexture(gcb_ctx* ctx, gcb_pitched_ptr p) {
    CTX(ctx);
    if (!C->d_tex) return fail_msg(C, "updateTexture: setupTexture not called");
    cudaMemcpy3DParms prm;
    memset(&prm, 0, sizeof prm);
    prm.srcPtr = make_cudaPitchedPtr(p.ptr, p.pitch, p.xsize, p.ysize);
    prm.dstPtr = make_cudaPitchedPtr(C->d_tex, (size_t)C->tex_x * sizeof(float), (size_t)C->tex_x * sizeof(float), C->tex_y);
    prm.extent = make_cudaExtent((size_t)C->tex_x * sizeof(float), C->tex_y, C->tex_z);
    prm.kind = cudaMemcpyDeviceToDevice;
    GCB_CHECK(C, cudaMemcpy3DAsync(&prm, C->stream));
    return field_call_end(C);
}
int gcb_deleteTexture(gcb_ctx* ctx) {
    CTX(ctx);
    if (C->d_tex) { cudaFree(C->d_tex); C->d_tex = nullptr; }
    return 0;
}

int gcb_file_write_obj(gcb_ctx* ctx, void* d_pos, unsigned int totalVerts, const char* filename) {
    CTX(ctx);
    // default: weld, face filter and text formatting on the GPU (obj_gpu.cu); same bytes as the host restatement below
    if (!(C->options & GCB_OPT_OBJ_HOST) && totalVerts < 0x7fffffffu) return write_obj_device(C, (const float4*)d_pos, totalVerts, filename);
    float* h = nullptr;
    if (totalVerts) {
        GCB_CHECK(C, cudaMallocHost(&h, (size_t)totalVerts * 16));
        cudaError_t e = cudaMemcpyAsync(h, d_pos, (size_t)totalVerts * 16, cudaMemcpyDeviceToHost, C->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(C->stream);
        if (e != cudaSuccess) { cudaFreeHost(h); return fail(C, "file_write_obj D2H", e); }
    }
    int r = write_obj_host(h, totalVerts, filename);
    if (h) cudaFreeHost(h);
    if (r) return fail_msg(C, std::string("file_write_obj: cannot write ") + filename);
    return 0;
}

// ------------------------------------------------------------------ SVL phase solve
int gcb_finding_phi(gcb_ctx* ctx, float* d_phi, float* d_period, int x_dim, int y_dim, int z_dim, int i, int j, int k, float dx, float dy, float dz,
                    char latticetype_one, int unform_type, float const_peirod, float x_period, float y_period, float z_period, float lcon, float lcon_1,
                    int sinewave_zaxis) {
    CTX(ctx);
    const int ijk[3] = {i, j, k};
    return k_finding_phi(C, d_phi, d_period, ijk, 1, x_dim, y_dim, z_dim, dx, dy, dz, latticetype_one, unform_type, const_peirod, x_period, y_period, z_period, lcon,
                         lcon_1, sinewave_zaxis);
}
int gcb_GPUCG_lattice(gcb_ctx* ctx, float* d_phi, int NX, int NY, int NZ, int iter, int OptIter, float EndRes, int* FinalIter, float* FinalRes) {
    CTX(ctx);
    (void)OptIter;
    return k_cg_batched(C, d_phi, 1, NX, NY, NZ, iter, EndRes, FinalIter, FinalRes);
}
int gcb_svl_phase_solve(gcb_ctx* ctx, float* d_phi_all, float* d_period, int nharm, const int* ijk, int x_dim, int y_dim, int z_dim, float dx, float dy, float dz,
                        char latticetype_one, int unform_type, float const_peirod, float x_period, float y_period, float z_period, float lcon, float lcon_1,
                        int sinewave_zaxis, int iter, float EndRes, int* FinalIter, float* FinalRes) {
    CTX(ctx);
    if (int r = k_finding_phi(C, d_phi_all, d_period, ijk, nharm, x_dim, y_dim, z_dim, dx, dy, dz, latticetype_one, unform_type, const_peirod, x_period, y_period,
                              z_period, lcon, lcon_1, sinewave_zaxis))
        return r;
    return k_cg_batched(C, d_phi_all, nharm, x_dim, y_dim, z_dim, iter, EndRes, FinalIter, FinalRes);
}

// ------------------------------------------------------------------ fused entry points
int gcb_svl_field(gcb_ctx* ctx, float* d_svl, const float* d_phi, int nh, const float* coef_host, int cx, int cy, int cz_local, int cz0, int NX2, int NY2,
                  int NZ2_local, gcb_slab slab, float dx, float dy, float dz, int accumulate, float* d_minmax) {
    CTX(ctx);
    if (d_minmax) if (int r = k_minmax_init(C, C->d_minmax)) return r;
    if (C->timing) cudaEventRecord(C->ev[2], C->stream);
    if (int r = k_svl_field(C, d_svl, d_phi, nh, coef_host, cx, cy, cz_local, cz0, NX2, NY2, NZ2_local, slab.z0, dx, dy, dz, accumulate,
                            d_minmax ? C->d_minmax : nullptr))
        return r;
    if (C->timing) { cudaEventRecord(C->ev[3], C->stream); C->field_timed = true; }
    if (d_minmax) if (int r = k_minmax_decode(C, C->d_minmax, d_minmax)) return r;
    return 0;
}

int gcb_internal_extract_band_raw(Ctx* C, const float* d_field, float a, float b, const float* d_ab, float isoValue, float isovalue1, float isovalue2,
                                 gcb_uint3 gridSizeLocal, gcb_slab slab, gcb_float3 voxelSize, gcb_float3 gridcenter, void* pos, void* norm,
                                 unsigned long long maxVerts, unsigned int* d_compVoxelArray, int count_only, unsigned long long* activeVoxels,
                                 unsigned long long* totalVerts, unsigned long long* h_totals_async) {
    if (!d_field) return fail_msg(C, "extract_band_raw: null field");
    McArgs A;
    base_args(A, M_BAND_RAW, gridSizeLocal, voxelSize, gridcenter, isoValue);
    A.iso1 = isovalue1; A.iso2 = isovalue2;
    A.f0 = d_field;
    A.na = a; A.nb = b;
    A.d_ab = d_ab;
    A.gz0 = slab.z0;
    A.gnz = slab.gnz ? slab.gnz : gridSizeLocal.z;
    A.pos = (float4*)pos; A.norm = (float4*)norm;
    A.max_verts = maxVerts;
    A.comp = d_compVoxelArray;
    A.count_only = count_only;
    unsigned long long act = 0, verts = 0;
    if (int r = launch_extract(C, A, &act, &verts, h_totals_async)) return r;
    if (h_totals_async) return 0;
    if (activeVoxels) *activeVoxels = act;
    if (totalVerts) *totalVerts = count_only ? C->h_totals[1] : verts;
    return 0;
}
int gcb_extract_band_raw(gcb_ctx* ctx, const float* d_field, float a, float b, float isoValue, float isovalue1, float isovalue2, gcb_uint3 gridSizeLocal,
                         gcb_slab slab, gcb_float3 voxelSize, gcb_float3 gridcenter, void* pos, void* norm, unsigned long long maxVerts,
                         unsigned int* d_compVoxelArray, int count_only, unsigned long long* activeVoxels, unsigned long long* totalVerts) {
    CTX(ctx);
    return gcb_internal_extract_band_raw(C, d_field, a, b, nullptr, isoValue, isovalue1, isovalue2, gridSizeLocal, slab, voxelSize, gridcenter, pos, norm, maxVerts,
                                 d_compVoxelArray, count_only, activeVoxels, totalVerts, nullptr);
}
int gcb_extract_band_raw_dev(gcb_ctx* ctx, const float* d_field, const float* d_minmax, float isoValue, float isovalue1, float isovalue2,
                             gcb_uint3 gridSizeLocal, gcb_slab slab, gcb_float3 voxelSize, gcb_float3 gridcenter, void* pos, void* norm,
                             unsigned long long maxVerts, unsigned int* d_compVoxelArray, int count_only, unsigned long long* activeVoxels,
                             unsigned long long* totalVerts) {
    CTX(ctx);
    if (!d_minmax) return fail_msg(C, "extract_band_raw_dev: null min/max pointer");
    return gcb_internal_extract_band_raw(C, d_field, 0.f, 0.f, d_minmax, isoValue, isovalue1, isovalue2, gridSizeLocal, slab, voxelSize, gridcenter, pos, norm, maxVerts,
                                 d_compVoxelArray, count_only, activeVoxels, totalVerts, nullptr);
}

int gcb_svl_lattice(gcb_ctx* ctx, float* d_svl_scratch, const float* d_phi, int nh, const float* coef_host, int cx, int cy, int cz, int NX2, int NY2, int NZ2,
                    float dx, float dy, float dz, float isoValue, float isovalue1, float isovalue2, gcb_float3 voxelSize, gcb_float3 gridcenter, void* pos,
                    void* norm, unsigned long long maxVerts, unsigned long long* activeVoxels, unsigned long long* totalVerts, float* minmax_out) {
    CTX(ctx);
    gcb_slab slab{0u, (unsigned)NZ2};
    // field -> min/max stays in device memory -> extraction reads it there: ONE synchronisation, at the end, for the counts
    if (int r = gcb_svl_field(ctx, d_svl_scratch, d_phi, nh, coef_host, cx, cy, cz, 0, NX2, NY2, NZ2, slab, dx, dy, dz, 0, C->d_minmax)) return r;
    if (minmax_out) GCB_CHECK(C, cudaMemcpyAsync(C->h_minmax, C->d_minmax, 2 * sizeof(float), cudaMemcpyDeviceToHost, C->stream));
    gcb_uint3 gs{(unsigned)NX2, (unsigned)NY2, (unsigned)NZ2};
    if (int r = gcb_internal_extract_band_raw(C, d_svl_scratch, 0.f, 0.f, C->d_minmax, isoValue, isovalue1, isovalue2, gs, slab, voxelSize, gridcenter, pos, norm, maxVerts,
                                      nullptr, 0, activeVoxels, totalVerts, nullptr))
        return r;
    if (minmax_out) { minmax_out[0] = C->h_minmax[0]; minmax_out[1] = C->h_minmax[1]; }
    return 0;
}

// d_mm: device scratch of the reduction (2 words, decoded in place to {min, max} floats), null = no reduction; d_minmax_out: where
// the caller wants the decoded pair (may equal d_mm or be null); scratch_free: event after which d_phi_scratch may be overwritten
// (null: everything queued on the compute stream so far)
static int svl_field_host_impl(gcb_ctx* ctx, Ctx* C, float* d_svl, const float* h_phi, float* d_phi_scratch, int nh, const float* coef_host, int cx, int cy,
                               int cz_local, int cz0, int NX2, int NY2, int NZ2_local, gcb_slab slab, float dx, float dy, float dz, float* d_mm,
                               float* d_minmax_out, cudaEvent_t scratch_free) {
    float* const d_minmax = d_mm;
    // Upload / compute overlap by z-slabs of the FINE grid: a slab of fine layers needs only the control planes that bracket
    // it, for all harmonics -- one strided copy (nh rows of `planes * cy * cx` floats) on a copy stream -- and is then evaluated
    // for all harmonics in one launch, so every tile pays its staging prologue once and the running sum never round-trips
    // through memory (an earlier version split the HARMONICS into batches instead: 8 launches over the whole volume, 8
    // prologues per tile and an accumulate pass each).  Thin slabs first: the first kernel starts after a short copy.
    if (!C->copy_stream) {
        GCB_CHECK(C, cudaStreamCreateWithFlags(&C->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < Ctx::kBatches; ++i) GCB_CHECK(C, cudaEventCreateWithFlags(&C->copy_ev[i], cudaEventDisableTiming));
    }
    if (!C->aux_stream) {
        GCB_CHECK(C, cudaStreamCreateWithFlags(&C->aux_stream, cudaStreamNonBlocking));
        GCB_CHECK(C, cudaEventCreateWithFlags(&C->aux_ev[0], cudaEventDisableTiming));
        GCB_CHECK(C, cudaEventCreateWithFlags(&C->aux_ev[1], cudaEventDisableTiming));
    }
    if (nh <= 0 || NZ2_local <= 0) return gcb_svl_field(ctx, d_svl, d_phi_scratch, nh, coef_host, cx, cy, cz_local, cz0, NX2, NY2, NZ2_local, slab, dx, dy, dz, 0, d_minmax_out);
    const size_t plane = (size_t)cx * cy, per = plane * cz_local;
    // slab thicknesses (fine layers, multiples of 8): 8, 8, 16, then ~1/14 of the grid each
    int f_start[Ctx::kBatches + 1];
    int nb = 0;
    {
        const int big = std::max(8, ((NZ2_local / 14) + 7) & ~7);
        int f = 0;
        const int lead[3] = {8, 8, 16};
        while (f < NZ2_local && nb < Ctx::kBatches) {
            f_start[nb] = f;
            const int left = Ctx::kBatches - nb;
            int t = nb < 3 ? std::min(lead[nb], big) : big;
            if (left == 1) t = NZ2_local - f;
            f = std::min(NZ2_local, f + t);
            ++nb;
        }
        f_start[nb] = NZ2_local;
    }
    // order the copy stream after the last reader of the scratch: an event of the caller's (job pipeline), else everything
    // already queued on the compute stream
    if (scratch_free) GCB_CHECK(C, cudaStreamWaitEvent(C->copy_stream, scratch_free, 0));
    else {
        GCB_CHECK(C, cudaEventRecord(C->copy_ev[0], C->stream));
        GCB_CHECK(C, cudaStreamWaitEvent(C->copy_stream, C->copy_ev[0], 0));
    }
    int copied = 0;  // local control planes [0, copied) are already queued
    for (int b = 0; b < nb; ++b) {
        // control planes bracketing fine layers [f_start[b], f_start[b+1]) (+1 plane of slack against rounding in generic ratios)
        const double zhi = ((double)f_start[b + 1] - 1.0 + (double)slab.z0) * (double)dz;
        int p_hi = (int)floor(zhi) + 2 - cz0;
        p_hi = std::min(std::max(p_hi, 0), cz_local - 1);
        if (b == nb - 1) p_hi = cz_local - 1;
        if (p_hi + 1 > copied) {
            const size_t off = (size_t)copied * plane, width = (size_t)(p_hi + 1 - copied) * plane * sizeof(float);
            GCB_CHECK(C, cudaMemcpy2DAsync(d_phi_scratch + off, per * sizeof(float), h_phi + off, per * sizeof(float), width, (size_t)nh, cudaMemcpyHostToDevice,
                                           C->copy_stream));
            copied = p_hi + 1;
        }
        GCB_CHECK(C, cudaEventRecord(C->copy_ev[b], C->copy_stream));
    }
    if (d_minmax) if (int r = k_minmax_init(C, d_mm)) return r;
    if (C->timing) cudaEventRecord(C->ev[2], C->stream);
    static const bool trace = getenv("GCB_TRACE") != nullptr;  // debugging aid: per-slab timeline on stderr
    cudaEvent_t tev[Ctx::kBatches + 1];
    if (trace) { for (int b = 0; b <= nb; ++b) cudaEventCreate(&tev[b]); cudaEventRecord(tev[0], C->stream); }
    // slabs alternate between the context's stream and a second one: they write disjoint layers (min/max goes through atomics),
    // so the last, partly filled wave of one launch overlaps the first waves of the next
    GCB_CHECK(C, cudaEventRecord(C->aux_ev[0], C->stream));              // min/max init precedes the aux launches
    GCB_CHECK(C, cudaStreamWaitEvent(C->aux_stream, C->aux_ev[0], 0));
    cudaStream_t main_stream = C->stream;
    for (int b = 0; b < nb; ++b) {
        cudaStream_t st = (b & 1) ? C->aux_stream : main_stream;
        GCB_CHECK(C, cudaStreamWaitEvent(st, C->copy_ev[b], 0));
        const int f0 = f_start[b], f1 = f_start[b + 1];
        C->stream = st;
        const int r = k_svl_field(C, d_svl + (size_t)f0 * NX2 * NY2, d_phi_scratch, nh, coef_host, cx, cy, cz_local, cz0, NX2, NY2, f1 - f0, slab.z0 + (unsigned)f0, dx,
                                  dy, dz, 0, d_mm);
        C->stream = main_stream;
        if (r) return r;
        if (trace) cudaEventRecord(tev[b + 1], st);
    }
    GCB_CHECK(C, cudaEventRecord(C->aux_ev[1], C->aux_stream));
    GCB_CHECK(C, cudaStreamWaitEvent(C->stream, C->aux_ev[1], 0));
    if (trace) {
        cudaStreamSynchronize(C->stream);
        for (int b = 0; b < nb; ++b) { float ms; cudaEventElapsedTime(&ms, tev[0], tev[b + 1]); fprintf(stderr, "[gcb trace] slab %d (fine layers %d..%d) done at %.3f ms\n", b, f_start[b], f_start[b + 1], ms); }
        for (int b = 0; b <= nb; ++b) cudaEventDestroy(tev[b]);
    }
    if (C->timing) { cudaEventRecord(C->ev[3], C->stream); C->field_timed = true; }
    if (d_mm) if (int r = k_minmax_decode(C, d_mm, d_minmax_out ? d_minmax_out : d_mm)) return r;
    return 0;
}
int gcb_svl_field_host(gcb_ctx* ctx, float* d_svl, const float* h_phi, float* d_phi_scratch, int nh, const float* coef_host, int cx, int cy, int cz_local,
                       int cz0, int NX2, int NY2, int NZ2_local, gcb_slab slab, float dx, float dy, float dz, float* d_minmax) {
    CTX(ctx);
    return svl_field_host_impl(ctx, C, d_svl, h_phi, d_phi_scratch, nh, coef_host, cx, cy, cz_local, cz0, NX2, NY2, NZ2_local, slab, dx, dy, dz,
                               d_minmax ? C->d_minmax : nullptr, d_minmax, nullptr);
}

int gcb_svl_lattice_host(gcb_ctx* ctx, const float* h_phi, float* d_phi_scratch, float* d_svl_scratch, int nh, const float* coef_host, int cx, int cy, int cz,
                         int NX2, int NY2, int NZ2, float dx, float dy, float dz, float isoValue, float isovalue1, float isovalue2, gcb_float3 voxelSize,
                         gcb_float3 gridcenter, void* pos, void* norm, unsigned long long maxVerts, unsigned long long* activeVoxels,
                         unsigned long long* totalVerts, float* minmax_out) {
    CTX(ctx);
    gcb_slab slab{0u, (unsigned)NZ2};
    if (int r = svl_field_host_impl(ctx, C, d_svl_scratch, h_phi, d_phi_scratch, nh, coef_host, cx, cy, cz, 0, NX2, NY2, NZ2, slab, dx, dy, dz, C->d_minmax, C->d_minmax,
                                    nullptr))
        return r;
    if (minmax_out) GCB_CHECK(C, cudaMemcpyAsync(C->h_minmax, C->d_minmax, 2 * sizeof(float), cudaMemcpyDeviceToHost, C->stream));
    gcb_uint3 gs{(unsigned)NX2, (unsigned)NY2, (unsigned)NZ2};
    if (int r = gcb_internal_extract_band_raw(C, d_svl_scratch, 0.f, 0.f, C->d_minmax, isoValue, isovalue1, isovalue2, gs, slab, voxelSize, gridcenter, pos, norm, maxVerts,
                                      nullptr, 0, activeVoxels, totalVerts, nullptr))
        return r;
    if (minmax_out) { minmax_out[0] = C->h_minmax[0]; minmax_out[1] = C->h_minmax[1]; }
    return 0;
}

// ------------------------------------------------------------------ fused unit-lattice path (BASELINE config 1)
// Multitopo::display_unit_lattice (main.cu:4113-4132): create_lattice -> GPU_buffer_normalise_buffer -> GPU_buffer_normalise_four ->
// computeIsosurface_latticeone, four blocking calls with two full min/max passes and three intermediate fields.  Here: the raw
// field once (with its TRUE range reduced on the fly), four derived range numbers in device memory, and the band-raw extraction
// applying both normalisations to the staged values -- one synchronisation, for the counts.
static int band_lattice_two_stage(Ctx* C, const float* d_field, gcb_uint3 gridSize, float isoValue, float isovalue1, float isovalue2, gcb_float3 voxelSize,
                                  gcb_float3 gridcenter, void* pos, void* norm, unsigned long long maxVerts, unsigned int* d_comp,
                                  unsigned long long* activeVoxels, unsigned long long* totalVerts, float* ranges_out) {
    unsigned* tmm = reinterpret_cast<unsigned*>(C->d_range4);
    float* ab4 = C->d_range4 + 2;
    if (int r = k_two_stage_range(C, tmm, ab4)) return r;
    if (ranges_out) GCB_CHECK(C, cudaMemcpyAsync(C->h_minmax, ab4, 4 * sizeof(float), cudaMemcpyDeviceToHost, C->stream));
    McArgs A;
    base_args(A, M_BAND_RAW, gridSize, voxelSize, gridcenter, isoValue);
    A.iso1 = isovalue1; A.iso2 = isovalue2;
    A.f0 = d_field;
    A.d_ab = ab4;
    A.two_stage = 1;
    A.gz0 = 0; A.gnz = gridSize.z;
    A.pos = (float4*)pos; A.norm = (float4*)norm;
    A.max_verts = maxVerts;
    A.comp = d_comp;
    unsigned long long act = 0, verts = 0;
    if (int r = launch_extract(C, A, &act, &verts)) return r;
    if (activeVoxels) *activeVoxels = act;
    if (totalVerts) *totalVerts = verts;
    if (ranges_out) for (int i = 0; i < 4; ++i) ranges_out[i] = C->h_minmax[i];
    return 0;
}
static int range4_init(Ctx* C) {
    if (!C->d_range4) GCB_CHECK(C, cudaMalloc(&C->d_range4, 32));
    return 0;
}
int gcb_band_lattice_from_raw(gcb_ctx* ctx, const float* d_raw_field, gcb_uint3 gridSize, float isoValue, float isovalue1, float isovalue2, gcb_float3 voxelSize,
                              gcb_float3 gridcenter, void* pos, void* norm, unsigned long long maxVerts, unsigned int* d_compVoxelArray,
                              unsigned long long* activeVoxels, unsigned long long* totalVerts, float* ranges_out) {
    CTX(ctx);
    if (!d_raw_field) return fail_msg(C, "band_lattice_from_raw: null field");
    if (int r = range4_init(C)) return r;
    if (int r = k_true_minmax(C, d_raw_field, (size_t)gridSize.x * gridSize.y * gridSize.z, reinterpret_cast<unsigned*>(C->d_range4))) return r;
    return band_lattice_two_stage(C, d_raw_field, gridSize, isoValue, isovalue1, isovalue2, voxelSize, gridcenter, pos, norm, maxVerts, d_compVoxelArray, activeVoxels,
                                  totalVerts, ranges_out);
}
int gcb_tpms_lattice(gcb_ctx* ctx, float* d_field_scratch, unsigned int lattice_type_index, gcb_uint3 gridSize, float isoValue, float isovalue1, float isovalue2,
                     gcb_float3 voxelSize, gcb_float3 gridcenter, void* pos, void* norm, unsigned long long maxVerts, unsigned int* d_compVoxelArray,
                     unsigned long long* activeVoxels, unsigned long long* totalVerts, float* ranges_out) {
    CTX(ctx);
    if (!d_field_scratch) return fail_msg(C, "tpms_lattice: null field scratch");
    if (lattice_type_index > 5u) return fail_msg(C, "tpms_lattice: lattice type must be 0..5");
    if (int r = range4_init(C)) return r;
    if (int r = k_create_lattice(C, d_field_scratch, gridSize.x, gridSize.y, gridSize.z, lattice_type_index, reinterpret_cast<unsigned*>(C->d_range4))) return r;
    return band_lattice_two_stage(C, d_field_scratch, gridSize, isoValue, isovalue1, isovalue2, voxelSize, gridcenter, pos, norm, maxVerts, d_compVoxelArray,
                                  activeVoxels, totalVerts, ranges_out);
}

// ------------------------------------------------------------------ fused density-surface path (BASELINE config 5)
// Multitopo::toprun (main.cu:3060-3109): copytotexture + updateTexture + refine (2x trilinear upsample of the coarse density) +
// computeIsosurface_2 with an all-zero vol_topo / d_result.  Here the upsample reads the caller's coarse grid where it lies (no
// staging copies) and the extraction skips the 16-byte grid_points stream and the colour field it would read as zeros.
int gcb_density_surface(gcb_ctx* ctx, const float* d_coarse, int cx, int cy, int cz, float* d_density_fine, int NX2, int NY2, int NZ2, float dx, float dy, float dz,
                        float isoValue, gcb_float3 voxelSize, gcb_float3 gridcenter, void* pos, void* norm, unsigned long long maxVerts,
                        unsigned int* d_compVoxelArray, unsigned long long* activeVoxels, unsigned long long* totalVerts) {
    CTX(ctx);
    if (!d_coarse || !d_density_fine) return fail_msg(C, "density_surface: null density buffer");
    if (int r = k_refine(C, d_coarse, cx, cy, cz, d_density_fine, NX2, NY2, NZ2, dx, dy, dz)) return r;
    McArgs A;
    gcb_uint3 gs{(unsigned)NX2, (unsigned)NY2, (unsigned)NZ2};
    base_args(A, M_TOPO, gs, voxelSize, gridcenter, isoValue);
    A.iso1 = 0.f;           // isovalue1 of the reference call; vol_topo.val == 0 never passes `val < 0`
    A.f0 = d_density_fine;
    A.gz0 = 0; A.gnz = gs.z;
    A.pos = (float4*)pos; A.norm = (float4*)norm;
    A.max_verts = maxVerts;
    A.comp = d_compVoxelArray;
    unsigned long long act = 0, verts = 0;
    if (int r = launch_extract(C, A, &act, &verts)) return r;
    if (activeVoxels) *activeVoxels = act;
    if (totalVerts) *totalVerts = verts;
    return 0;
}

// ------------------------------------------------------------------ two-deep job pipeline of the host-input path
static int slot_init(Ctx* C, Ctx::Slot& s) {
    if (s.h_totals) return 0;
    GCB_CHECK(C, cudaMallocHost(&s.h_totals, 16));
    GCB_CHECK(C, cudaMallocHost(&s.h_minmax, 16));
    GCB_CHECK(C, cudaMalloc(&s.d_minmax, 16));
    GCB_CHECK(C, cudaEventCreateWithFlags(&s.field_done, cudaEventDisableTiming));
    GCB_CHECK(C, cudaEventCreateWithFlags(&s.job_done, cudaEventDisableTiming));
    return 0;
}
int gcb_svl_lattice_host_submit(gcb_ctx* ctx, int slot, const float* h_phi, float* d_phi_scratch, float* d_svl_scratch, int nh, const float* coef_host, int cx,
                                int cy, int cz, int NX2, int NY2, int NZ2, float dx, float dy, float dz, float isoValue, float isovalue1, float isovalue2,
                                gcb_float3 voxelSize, gcb_float3 gridcenter, void* pos, void* norm, unsigned long long maxVerts) {
    CTX(ctx);
    if (slot < 0 || slot > 1) return fail_msg(C, "svl_lattice_host_submit: slot must be 0 or 1");
    Ctx::Slot& S = C->slot[slot];
    if (S.busy || S.field_pending) return fail_msg(C, "svl_lattice_host_submit: slot still holds an unfinished job (call gcb_svl_lattice_host_wait first)");
    if (int r = slot_init(C, S)) return r;
    gcb_slab slab{0u, (unsigned)NZ2};
    // the control grids of this job are copied while the previous job (other slot, other scratch) computes: the copy stream only
    // waits for the last field kernel that read THIS slot's scratch
    if (int r = svl_field_host_impl(ctx, C, d_svl_scratch, h_phi, d_phi_scratch, nh, coef_host, cx, cy, cz, 0, NX2, NY2, NZ2, slab, dx, dy, dz, S.d_minmax, S.d_minmax,
                                    S.field_done))
        return r;
    GCB_CHECK(C, cudaEventRecord(S.field_done, C->stream));
    GCB_CHECK(C, cudaMemcpyAsync(S.h_minmax, S.d_minmax, 2 * sizeof(float), cudaMemcpyDeviceToHost, C->stream));
    gcb_uint3 gs{(unsigned)NX2, (unsigned)NY2, (unsigned)NZ2};
    if (int r = gcb_internal_extract_band_raw(C, d_svl_scratch, 0.f, 0.f, S.d_minmax, isoValue, isovalue1, isovalue2, gs, slab, voxelSize, gridcenter, pos, norm, maxVerts,
                                      nullptr, 0, nullptr, nullptr, S.h_totals))
        return r;
    GCB_CHECK(C, cudaEventRecord(S.job_done, C->stream));
    S.busy = true;
    return 0;
}
// The same pipeline for one z-slab of a sharded job: the range of the whole field only exists after an exchange between the ranks, so the
// job is enqueued in two halves with the caller's (stream-ordered) reduction in between.
int gcb_svl_slab_host_submit_field(gcb_ctx* ctx, int slot, const float* h_phi, float* d_phi_scratch, float* d_svl_scratch, int nh, const float* coef_host, int cx,
                                   int cy, int cz_local, int cz0, int NX2, int NY2, int NZ2_local, gcb_slab slab, float dx, float dy, float dz, float* d_minmax) {
    CTX(ctx);
    if (slot < 0 || slot > 1) return fail_msg(C, "svl_slab_host_submit_field: slot must be 0 or 1");
    if (!d_minmax) return fail_msg(C, "svl_slab_host_submit_field: d_minmax (device float[2]) is required");
    Ctx::Slot& S = C->slot[slot];
    if (S.busy || S.field_pending) return fail_msg(C, "svl_slab_host_submit_field: slot still holds an unfinished job");
    if (int r = slot_init(C, S)) return r;
    if (int r = svl_field_host_impl(ctx, C, d_svl_scratch, h_phi, d_phi_scratch, nh, coef_host, cx, cy, cz_local, cz0, NX2, NY2, NZ2_local, slab, dx, dy, dz, S.d_minmax,
                                    d_minmax, S.field_done))
        return r;
    GCB_CHECK(C, cudaEventRecord(S.field_done, C->stream));
    S.field_pending = true;
    return 0;
}
int gcb_svl_slab_host_submit_extract(gcb_ctx* ctx, int slot, const float* d_svl_scratch, const float* d_ab, float isoValue, float isovalue1, float isovalue2,
                                     gcb_uint3 gridSizeLocal, gcb_slab slab, gcb_float3 voxelSize, gcb_float3 gridcenter, void* pos, void* norm,
                                     unsigned long long maxVerts) {
    CTX(ctx);
    if (slot < 0 || slot > 1) return fail_msg(C, "svl_slab_host_submit_extract: slot must be 0 or 1");
    if (!d_ab) return fail_msg(C, "svl_slab_host_submit_extract: d_ab (device float[2], the range over all ranks) is required");
    Ctx::Slot& S = C->slot[slot];
    if (!S.field_pending) return fail_msg(C, "svl_slab_host_submit_extract: no field submitted in this slot");
    if (int r = gcb_internal_extract_band_raw(C, d_svl_scratch, 0.f, 0.f, d_ab, isoValue, isovalue1, isovalue2, gridSizeLocal, slab, voxelSize, gridcenter, pos, norm,
                                              maxVerts, nullptr, 0, nullptr, nullptr, S.h_totals))
        return r;
    GCB_CHECK(C, cudaMemcpyAsync(S.h_minmax, d_ab, 2 * sizeof(float), cudaMemcpyDeviceToHost, C->stream));
    GCB_CHECK(C, cudaEventRecord(S.job_done, C->stream));
    S.field_pending = false;
    S.busy = true;
    return 0;
}
int gcb_svl_lattice_host_wait(gcb_ctx* ctx, int slot, unsigned long long* activeVoxels, unsigned long long* totalVerts, float* minmax_out) {
    CTX(ctx);
    if (slot < 0 || slot > 1) return fail_msg(C, "svl_lattice_host_wait: slot must be 0 or 1");
    Ctx::Slot& S = C->slot[slot];
    if (!S.busy) return fail_msg(C, "svl_lattice_host_wait: no job in this slot");
    GCB_CHECK(C, cudaEventSynchronize(S.job_done));
    S.busy = false;
    if (activeVoxels) *activeVoxels = S.h_totals[0];
    if (totalVerts) *totalVerts = S.h_totals[0] ? S.h_totals[1] : 0;
    if (minmax_out) { minmax_out[0] = S.h_minmax[0]; minmax_out[1] = S.h_minmax[1]; }
    return 0;
}

} // extern "C"
