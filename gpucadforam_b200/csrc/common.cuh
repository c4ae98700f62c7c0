// common.cuh -- shared internals of libgpucad_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/gpucad_b200.h"

namespace gcb {

// 16-byte AoS CSG state; reference src/MarchingCubes_kernel.h:12-18
struct __align__(16) GridPoint { int val; float t_x, t_y, t_z; };
static_assert(sizeof(GridPoint) == sizeof(gcb_grid_points), "layout");

enum Mode : int {
    M_LATTICE_ONE = 0,  // computeIsosurface_latticeone  (mask `vol` + k `vol_one`)
    M_LATTICE = 1,      // computeIsosurface_lattice     (+ vol_two, iso1/iso2)
    M_CSG = 2,          // computeIsosurface             (grid_points + dynamic + lattice)
    M_TOPO = 3,         // computeIsosurface_2 / _topo   (vol_topo + density + result [+disp])
    M_BAND_RAW = 4,     // fused: raw field -> normalise + domain faces + band mask -> latticeone
    M_REGION = 5        // computeIsosurface_region      (vol_topo + primitive_fixed + dynamic, triangle_metadata)
};
enum : uint32_t {
    F_UNION = 1, F_DIFF = 2, F_INTERSECT = 4, F_FIXED = 8, F_DYNAMIC = 16, F_MAKE_REGION = 32, F_DISP = 64,
    F_SHOW_REGION = 128, F_SHOW_DOMAIN = 256
};
// 64-byte per-triangle record of the region variant; reference src/MarchingCubes_kernel.h:20-32
struct TriangleMetadata {
    unsigned int index, voxel, l_index, edge_1, edge_2, edge_3, load_group;
    float centroid[3], normal[3], force_dir[3];
};
static_assert(sizeof(TriangleMetadata) == sizeof(gcb_triangle_metadata) && sizeof(TriangleMetadata) == 64, "layout");

// Arguments of the fused extraction kernel (by value, < 4 KB).
struct McArgs {
    int mode;
    uint32_t nx, ny, nz;   // grid POINTS held by this rank (slab incl. +z halo plane)
    uint32_t cx, cy, cz;   // cells
    float3 voxel, center;
    float iso, iso1, iso2, iso1b, iso2b;
    uint32_t flags;
    unsigned long long max_verts;
    const float* f0;       // TMA-staged interpolation field (k | dynamic | density | raw)
    const float* f1;       // mask (lattice) | lattice_field (CSG) | d_result (topo)
    const float* f2;       // vol_two (M_LATTICE)
    const GridPoint* gp;   // primitive_fixed (CSG, region) | vol_topo (topo)
    const GridPoint* gp2;  // vol_topo (region)
    TriangleMetadata* meta;  // region + show_region: one record per triangle
    const float4* disp;    // topo displaced positions
    const float* f0_top;   // stored-field sharding: local point layer nz - 1 of f0 / f1 / gp when it is held by the upper neighbour
    const float* f1_top;   //   (pointer to the start of that layer; null = the layer is part of the local arrays)
    const GridPoint* gp_top;
    unsigned long long top_begin;  // (nz - 1) * nx * ny
    float na, nb;          // M_BAND_RAW: k = (f - na) / (nb - na)
    const float* d_ab;     // M_BAND_RAW: {na, nb} in device memory (overrides na / nb when non-null)
    int two_stage;         // M_BAND_RAW with d_ab: d_ab[2..3] = {a2, b2}, second normalisation k <- (k - a2) / (b2 - a2)
    uint32_t gz0, gnz;     // global z offset of local point layer 0, global number of point layers
    float4* pos;
    float4* norm;
    uint32_t* comp;        // compacted active cell ids (global linear id), may be null
    uint32_t* st_verts;    // optional stage arrays (GCB_OPT_FILL_STAGE_ARRAYS)
    uint32_t* st_occ;
    uint32_t* st_verts_scan;
    uint32_t* st_occ_scan;
    unsigned long long* status_a;  // decoupled look-back: [flag:2 | active prefix:62] per tile
    unsigned long long* status_v;  // [flag:2 | vertex prefix:62] per tile
    uint32_t* tile_counter;
    unsigned long long* totals;    // [0]=active cells, [1]=vertices
    uint32_t rows_per_tile, tiles_per_slice, num_tiles;
    uint32_t prow_stride;          // floats between the two staged slices in smem ( (R+1)*nx rounded )
    int use_tma;
    int count_only;
    uint32_t ppr, cpr;                      // 128-point chunks per staged row, 128-cell chunks per cell row
    uint32_t cstride;                       // bytes between cell rows of the cube-index array in smem (128 * cpr)
    float snap_thr;                         // smallest float >= 0.0005 (endpoint snapping threshold)
};

struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    unsigned int options = GCB_OPT_LEGACY_MEMSET;
    std::string err;
    unsigned long long launches = 0;
    int num_sms = 148;
    // extraction scratch
    unsigned long long* d_status = nullptr;  // status_cap words: totals, tile counter, look-back state of both prefixes
    size_t status_cap = 0;
    uint32_t* d_tile_counter = nullptr;
    unsigned long long* d_totals = nullptr;  // 2 words
    unsigned long long* h_totals = nullptr;  // pinned
    // reductions
    float* d_minmax = nullptr;               // 2 floats (ordered-int encoded during reduction)
    float* d_tab = nullptr; size_t tab_cap = 0;  // separable-primitive tables (fields.cu line_tab_kernel)
    float* d_range4 = nullptr;               // fused normalise-twice paths: 2 raw words (true min / max) + {a, b, a2, b2}
    float* h_minmax = nullptr;               // pinned
    // control grid ("texture")
    float* d_tex = nullptr;
    int tex_x = 0, tex_y = 0, tex_z = 0;
    // SVL coefficient staging
    float* d_coef = nullptr; size_t coef_cap = 0;
    // legacy table copies
    unsigned int* d_tri = nullptr; unsigned int* d_nverts = nullptr;
    // timing
    bool timing = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool extract_timed = false, field_timed = false;
    // host-input pipeline: control grids uploaded in batches on a copy stream, overlapped with the field kernel
    static constexpr int kBatches = 20;
    cudaStream_t copy_stream = nullptr;
    cudaStream_t aux_stream = nullptr;   // second compute stream of the slab-overlapped host path (tail of slab k overlaps head of k+1)
    cudaEvent_t aux_ev[2] = {};
    cudaEvent_t copy_ev[kBatches] = {};
    // two-deep job pipeline of the host-input entry point (gcb_svl_lattice_host_submit / _wait): per slot the pinned result
    // words, the device min/max pair of the job's field, and an event marking the end of the job's last reader of the
    // caller's control-grid scratch
    struct Slot {
        unsigned long long* h_totals = nullptr;  // pinned: {active, vertices}
        float* h_minmax = nullptr;               // pinned: {min, max}
        float* d_minmax = nullptr;               // device: 2 words raw (ordered-int) + 2 floats decoded
        cudaEvent_t field_done = nullptr, job_done = nullptr;
        bool busy = false;           // a complete job is enqueued: _wait may be called
        bool field_pending = false;  // sharded job: the field half is enqueued, the extraction half is not yet
    } slot[2];
};

int fail(Ctx* c, const char* what, cudaError_t e);
int fail_msg(Ctx* c, const std::string& msg);
#define GCB_CHECK(c, call)                                          \
    do {                                                            \
        cudaError_t e_ = (call);                                    \
        if (e_ != cudaSuccess) return gcb::fail((c), #call, e_);    \
    } while (0)

static inline float3 f3(gcb_float3 v) { return make_float3(v.x, v.y, v.z); }
inline void base_args(McArgs& a, int mode, gcb_uint3 gridSize, gcb_float3 voxelSize, gcb_float3 gridcenter, float iso) {
    memset(&a, 0, sizeof a);
    a.mode = mode;
    a.nx = gridSize.x; a.ny = gridSize.y; a.nz = gridSize.z;
    a.voxel = f3(voxelSize);
    a.center = f3(gridcenter);
    a.iso = iso;
}

// ---- launchers implemented in the .cu files ----
// h_totals_async != null: enqueue only (kernel + D2H of {active, vertices} into that pinned slot), no synchronisation
int launch_extract(Ctx* c, McArgs& a, unsigned long long* active, unsigned long long* verts, unsigned long long* h_totals_async = nullptr);
void host_tables(unsigned int* tri, unsigned int* nverts);
int upload_tables_legacy(Ctx* c);

// fields.cu
int k_create_lattice(Ctx* c, float* out, unsigned nx, unsigned ny, unsigned nz, unsigned type, unsigned* d_true_minmax = nullptr);
int k_true_minmax(Ctx* c, const float* in, size_t n, unsigned* d_true_minmax);
int k_two_stage_range(Ctx* c, const unsigned* d_true_minmax, float* d_ab4);
int k_unit_spectrum(Ctx* c, const float* f, int nx, int ny, int nz, int range, float2* out);
int k_sphere(Ctx* c, float* out, float3 center, float radius, float thickness, int nx, int ny, int nz, float dx, float dy, float dz, bool shell);
int k_line(Ctx* c, float* out, float3 center, float3 axis, float radius, float tr, float ta, int nx, int ny, int nz, float dx, float dy, float dz, bool disc);
int k_cuboid(Ctx* c, float* out, float3 center, float3 ang, float xw, float yw, float zw, int nx, int ny, int nz, float dx, float dy, float dz);
int k_cuboid_shell(Ctx* c, float* out, float3 center, float3 ang, float xw, float yw, float zw, float th, int nx, int ny, int nz, float dx, float dy, float dz);
int k_torus(Ctx* c, float* out, float3 center, float3 ang, float R, float rc, int nx, int ny, int nz, float dx, float dy, float dz);
int k_cone(Ctx* c, float* out, float3 center, float3 ang, float br, float h, int nx, int ny, int nz, float dx, float dy, float dz);
int k_cone_frustum(Ctx* c, float* out, float3 center, float3 ang, float tr, float br, float h, int nx, int ny, int nz, float dx, float dy, float dz);
int k_pyramid_frustum(Ctx* c, float* out, float3 center, float3 ang, float xb, float xt, float yh, float zb, float zt, int nx, int ny, int nz, float dx, float dy, float dz);
int k_minmax(Ctx* c, const float* in, size_t n, float* lo, float* hi);           // host results, reference semantics
int k_minmax_device(Ctx* c, const float* in, size_t n);                          // leaves decoded result in c->d_minmax
int k_normalise(Ctx* c, const float* in, float* out, size_t n, float a, float b, const float* d_ab = nullptr);
int k_normalise_four(Ctx* c, const float* in, float* mask, float* k, int nx, int ny, int nz, float a, float b, float iso1, float iso2,
                     const float* d_ab = nullptr);
int k_refine(Ctx* c, const float* tex, int cx, int cy, int cz, float* out, int nx2, int ny2, int nz2, float dx, float dy, float dz);
int k_grating(Ctx* c, const float* tex, int cx, int cy, int cz, float2* out, int nx2, int ny2, int nz2, float dx, float dy, float dz);
int k_svl(Ctx* c, float* svl, const float2* grating, size_t n, int idx, const float2* coef);
int k_svl_field(Ctx* c, float* svl, const float* phi, int nh, const float* d_coef, int cx, int cy, int czl, int cz0, int nx2, int ny2, int nz2l,
                unsigned z0, float dx, float dy, float dz, int accumulate, float* d_minmax_raw);
int k_minmax_init(Ctx* c, float* d_minmax_raw);
int k_minmax_decode(Ctx* c, float* d_minmax_raw, float* d_out);
int k_csg_retain_primitive(Ctx* c, int kind, float3 center, float3 aux, const float* params, int nparams, int flag, float* d_field, GridPoint* vol_one, int nx,
                           int ny, int nz, float dx, float dy, float dz, float iso, bool u, bool d, bool i);
int k_copy_parameter(Ctx* c, GridPoint* vol_one, const float* vol_two, const float* vol_lattice, bool dynamic, float iso1, float iso2,
                     unsigned nx, unsigned ny, unsigned nz, float iso, bool u, bool d, bool i);
int k_primitive_field(Ctx* c, const GridPoint* prim, const float* active, float* isosurf, size_t n, bool fixed, bool dynamic);
int k_topo_field(Ctx* c, const float* topo, float* isosurf, float volfrac, size_t n);
int k_patch_topo_field(Ctx* c, float* d, int nx, int ny, int nz, const GridPoint* vol_one);
int k_period_angle(Ctx* c, float* out, int nx, int ny, int nz, float dx, float dy, float dz, float mx, float my, float mz, int axis, bool angle);
int k_normalise_three(Ctx* c, const float* in, float* out, size_t n, float a1, float b1, const float* d_ab);
int k_copy_to_pitched(Ctx* c, const float* src, gcb_pitched_ptr dst, int nx, int ny, int nz);

// ---- linear point index -> (x, y, z).  The reference kernels do this with 64-bit / and % per point (up to four ~70-instruction
// division sequences, more than the rest of a primitive's arithmetic); the indices are exact integers either way, so here
// grids below 2^32 points use two multiply-shift divisions with host-computed magic numbers (Granlund-Montgomery, exact for
// every 32-bit numerator), larger grids the 64-bit form.
struct FastDiv { uint32_t d, m, s; };
static inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f{d, 0u, 0u};
    if (d > 1u) {
        uint32_t s = 0;
        while ((1ull << s) < d) ++s;
        f.s = s;
        f.m = (uint32_t)((((1ull << 32) * ((1ull << s) - d)) / d) + 1ull);
    }
    return f;
}
__device__ __forceinline__ uint32_t fast_div(uint32_t n, const FastDiv& f) {
    if (f.d == 1u) return n;
    const uint32_t t = __umulhi(n, f.m);
    return (t + ((n - t) >> 1)) >> (f.s - 1u);
}
struct Grid3 { uint32_t nx, ny; FastDiv by_nx, by_nxny; int small; };
static inline Grid3 make_grid3(size_t nx, size_t ny, size_t nz) {
    Grid3 g{(uint32_t)nx, (uint32_t)ny, make_fastdiv((uint32_t)nx), make_fastdiv((uint32_t)std::min<size_t>(nx * ny, 0xffffffffull)), 0};
    g.small = nx * ny * nz <= 0xffffffffull && nx * ny <= 0x7fffffffull;
    return g;
}
__device__ __forceinline__ void point_xyz(size_t i, const Grid3& g, int& x, int& y, int& z) {
    if (g.small) {
        const uint32_t n = (uint32_t)i, zz = fast_div(n, g.by_nxny), r = n - zz * g.by_nxny.d, yy = fast_div(r, g.by_nx);
        x = (int)(r - yy * g.nx); y = (int)yy; z = (int)zz;
    } else {
        z = (int)(i / ((size_t)g.nx * g.ny));
        y = (int)((i % ((size_t)g.nx * g.ny)) / g.nx);
        x = (int)(i % g.nx);
    }
}

// capi.cu: band-raw extraction with all knobs; h_totals_async != null = enqueue only (multi.cu)
extern "C" int gcb_internal_extract_band_raw(Ctx* C, const float* d_field, float a, float b, const float* d_ab, float isoValue, float isovalue1, float isovalue2,
                                             gcb_uint3 gridSizeLocal, gcb_slab slab, gcb_float3 voxelSize, gcb_float3 gridcenter, void* pos, void* norm,
                                             unsigned long long maxVerts, unsigned int* d_compVoxelArray, int count_only, unsigned long long* activeVoxels,
                                             unsigned long long* totalVerts, unsigned long long* h_totals_async);

// phase_solve.cu
int k_finding_phi(Ctx* c, float* phi_all, const float* period, const int* ijk_host, int nharm, int nx, int ny, int nz, float dx, float dy, float dz, int latticetype,
                  int uniform_type, float const_period, float x_period, float y_period, float z_period, float lcon, float lcon_1, int sinewave_zaxis);
int k_cg_batched(Ctx* c, float* phi_all, int nharm, int nx, int ny, int nz, int iter, float end_res, int* final_iter, float* final_res);

// obj_writer.cpp (host restatement, GCB_OPT_OBJ_HOST) and obj_gpu.cu (device weld + text, the default)
int write_obj_host(const float* pos4, unsigned int total_verts, const char* filename);
int write_obj_device(Ctx* c, const float4* pos, unsigned int total_verts, const char* filename);

} // namespace gcb
