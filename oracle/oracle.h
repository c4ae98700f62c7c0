/*
 * oracle.h -- C ABI of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This library is a plain C++/OpenMP restatement of the
 * reference's CUDA algorithm for the implicit-field + marching-cubes path (the reference has
 * no CPU implementation).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product library
 * (libgpucad_b200.so) never links, loads or calls anything declared here.
 *
 * Parity pinning: the reference ships no tests, golden vectors or fixtures (SURVEY.md
 * section 4), so this oracle is pinned against the reference's own CUDA kernels compiled
 * unmodified for sm_100a (oracle/_ref, built by oracle/Makefile) and run on the GPU box by
 * tests/test_gpu_reference_parity.py, plus the table invariants in tests/test_tables.py.
 */
#ifndef GPUCAD_ORACLE_H
#define GPUCAD_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 16-byte AoS CSG state per grid point: reference src/MarchingCubes_kernel.h:12-18 */
typedef struct orc_grid_point { int32_t val; float t_x, t_y, t_z; } orc_grid_point;

/* extraction modes; each names the reference host entry point it restates */
enum {
    ORC_MODE_LATTICE_ONE = 0, /* Isosurface::computeIsosurface_latticeone  Isosurface.cu:488-572 */
    ORC_MODE_LATTICE     = 1, /* Isosurface::computeIsosurface_lattice     Isosurface.cu:401-486 */
    ORC_MODE_CSG         = 2, /* Isosurface::computeIsosurface             Isosurface.cu:44-134  */
    ORC_MODE_TOPO        = 3, /* Isosurface::computeIsosurface_2 / _topo   Isosurface.cu:243-398 */
    ORC_MODE_REGION      = 5  /* Isosurface::computeIsosurface_region      Isosurface.cu:150-239 */
};

/* flag bits of orc_mc_params.flags (CSG mode) */
enum {
    ORC_F_UNION = 1, ORC_F_DIFF = 2, ORC_F_INTERSECT = 4,
    ORC_F_FIXED = 8, ORC_F_DYNAMIC = 16, ORC_F_MAKE_REGION = 32,
    ORC_F_DISP = 64, /* TOPO: positions interpolated from disp field (computeIsosurface_topo disp=true) */
    ORC_F_SHOW_REGION = 128, ORC_F_SHOW_DOMAIN = 256 /* REGION: show_region / show_domain (else ORC_F_MAKE_REGION) */
};

/* per-triangle record of the region variant: reference src/MarchingCubes_kernel.h:20-32 */
typedef struct orc_triangle_metadata {
    uint32_t index, voxel, l_index, edge_1, edge_2, edge_3, load_group;
    float centroid[3], normal[3], force_dir[3];
} orc_triangle_metadata;

typedef struct orc_mc_params {
    int32_t mode;
    uint32_t nx, ny, nz;        /* grid POINTS per axis; cells are (n-1) per axis           */
    float voxel[3];             /* voxelSize                                                */
    float center[3];            /* gridcenter                                               */
    float iso;                  /* isoValue                                                 */
    float iso1, iso2;           /* band / lattice thresholds (isovalue1,isovalue2 | iso1,iso2 of CSG) */
    float iso1b, iso2b;         /* LATTICE only: trailing iso1, iso2 of computeIsosurface_lattice */
    uint32_t flags;
    uint32_t max_verts;         /* capacity in vertices; writes at index >= max_verts-3 dropped */
    /* inputs (unused ones may be NULL) */
    const float* f0;            /* LATTICE*: mask `vol`;  CSG: primitive_dynamic;  TOPO: density vol_two */
    const float* f1;            /* LATTICE*: k `vol_one`; CSG: lattice_field;      TOPO: d_result        */
    const float* f2;            /* LATTICE : vol_two                                                   */
    const orc_grid_point* gp;   /* CSG: primitive_fixed;  TOPO: vol_topo                               */
    const float* disp;          /* TOPO+DISP: float4 per point                                         */
    const orc_grid_point* gp2;  /* REGION: vol_topo (gp = primitive_fixed, f0 = primitive_dynamic)     */
    orc_triangle_metadata* meta;/* REGION + SHOW_REGION: one record per triangle                       */
} orc_mc_params;

/* Marching-cubes tables (Bourke): tri is 256x16 with 255 terminators, nverts is 256. */
void orc_tables(uint32_t* tri, uint32_t* nverts);

/* Full extraction.  Stage arrays may be NULL when not wanted.  pos/norm are float4 arrays of
 * max_verts entries (caller-zeroed if it cares about the tail).  Returns 0. */
int orc_extract(const orc_mc_params* p,
                uint32_t* voxel_verts, uint32_t* voxel_occupied,
                uint32_t* voxel_verts_scan, uint32_t* voxel_occupied_scan,
                uint32_t* comp_voxel_array,
                float* pos, float* norm,
                uint32_t* active_voxels, uint32_t* total_verts);

/* Count-only pass (classification + sums), no stage arrays. */
int orc_count(const orc_mc_params* p, uint64_t* active_voxels, uint64_t* total_verts);

/* ---- field producers ---- */
void orc_create_lattice(float* out, uint32_t nx, uint32_t ny, uint32_t nz, uint32_t type);
/* SVL phase solve (Gratings.cu:100-417, :875-974): right-hand side of one harmonic; CG in place (rhs in, solution out) */
void orc_finding_phi(float* phi, const float* period, int nx, int ny, int nz, int i, int j, int k, float dx, float dy, float dz, int latticetype,
                     int uniform_type, float const_period, float x_period, float y_period, float z_period, float lcon, float lcon_1, int sinewave_zaxis);
void orc_cg(float* phi, int nx, int ny, int nz, int iter, float end_res, int* final_iter, float* final_res);
/* unit-cell spectrum (main.cu:3577-3706): (2*range+1)^3 lowest DFT coefficients / point count, order k, j, i; out = (re, im) pairs */
void orc_unit_spectrum(const float* f, int nx, int ny, int nz, int range, float* out);
void orc_sphere(float* out, const float center[3], float radius, float thickness,
                int nx, int ny, int nz, float dx, float dy, float dz, int shell);
void orc_distance_from_line(float* out, const float center[3], const float axis[3], float radius,
                            float thickness_radial, float thickness_axial,
                            int nx, int ny, int nz, float dx, float dy, float dz, int disc);
void orc_cuboid(float* out, const float center[3], const float angles[3], float xw, float yw, float zw,
                int nx, int ny, int nz, float dx, float dy, float dz);
void orc_cuboid_shell(float* out, const float center[3], const float angles[3], float xw, float yw,
                      float zw, float thickness, int nx, int ny, int nz, float dx, float dy, float dz);
void orc_torus(float* out, const float center[3], const float angles[3], float torus_radius,
               float circle_radius, int nx, int ny, int nz, float dx, float dy, float dz);
void orc_cone(float* out, const float center[3], const float angles[3], float base_radius,
              float height, int nx, int ny, int nz, float dx, float dy, float dz);
void orc_cone_frustum(float* out, const float center[3], const float angles[3], float top_radius,
                      float bottom_radius, float height, int nx, int ny, int nz,
                      float dx, float dy, float dz);
void orc_pyramid_frustum(float* out, const float center[3], const float angles[3], float xw_base,
                         float xw_top, float y_height, float zw_base, float zw_top,
                         int nx, int ny, int nz, float dx, float dy, float dz);

/* min/max as the reference's two-stage reduction defines it (includes its clamp-through-zero quirk) */
void orc_minmax(const float* f, size_t n, float* lo, float* hi);
void orc_normalise_buffer(const float* in, float* out, size_t n);
void orc_normalise_four(const float* in, float* mask, float* k, int nx, int ny, int nz,
                        float iso1, float iso2);
void orc_normalise_four_ab(const float* in, float* mask, float* k, int nx, int ny, int nz,
                           float iso1, float iso2, float a, float b);

/* control grid -> fine grid: software model of the trilinear texture fetch */
void orc_refine(const float* coarse, int cx, int cy, int cz, float* fine, int nx2, int ny2, int nz2,
                float dx, float dy, float dz);
/* one harmonic of the spatially-variant lattice: svl += cos(phi)*re - sin(phi)*im */
void orc_svl_accumulate(float* svl, const float* phi_coarse, int cx, int cy, int cz,
                        int nx2, int ny2, int nz2, float dx, float dy, float dz, float re, float im);

/* CSG retain and mask helpers */
void orc_copy_parameter(orc_grid_point* vol_one, const float* vol_two, const float* vol_lattice,
                        int dynamic, float iso1, float iso2, int nx, int ny, int nz, float iso,
                        int obj_union, int obj_diff, int obj_intersect);
void orc_primitive_field(const orc_grid_point* prim, const float* active, float* isosurf, size_t n,
                         int fixed, int dynamic);
void orc_topo_field(const float* topo, float* isosurf, float volfrac, size_t n);
void orc_patch_topo_field(float* d, int nx, int ny, int nz, const orc_grid_point* vol_one);

/* period / angle fields of the SVL phase solve: set_period_kernel / set_theta_kernel (Gratings.cu:775-853) and
 * GPU_buffer_normalise_three = min/max reduction + device_bufferthree (Gratings.cu:1071-1087, :1539-1572) */
void orc_period_data(float* out, int nx, int ny, int nz, float dx, float dy, float dz, float mx, float my, float mz, int axis);
void orc_angle_data(float* out, int nx, int ny, int nz, float dx, float dy, float dz, float mx, float my, float mz, int axis);
void orc_normalise_three(const float* in, float* out, size_t n, float a1, float b1);

/* .obj writer (File_output::file_write_obj); pos is float4[total_verts] on the host */
int orc_write_obj(const float* pos, uint32_t total_verts, const char* filename);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
