// headless_main.cpp -- runs the hot path without GLFW / ImGui / Vulkan.
//
// The reference has no headless mode: class Multitopo (src/main.cu) owns every buffer and drives the path from
// its frame loop.  This harness replays the same call sequences through the C++ host mirror (gpucad_host.hpp):
//   config 1  check_unit_lattice / display_unit_lattice   main.cu:4080-4137   gyroid unit cell, band extraction
//   config 2  show_model + retain                         main.cu:3304-3465   sphere U box - cylinder, .obj export :4695-4778
//   config 5  toprun_struct (extraction part)             main.cu:3060-3109   refine + computeIsosurface_2
// and the fused entry point for config 3 (spatial_lattice_run, main.cu:3904-4037).  Buffers are plain cudaMalloc
// allocations with the layouts of vulkan_create_lattice_buffers / initMC_two (main.cu:2125-2161, :2714-2814).
//
// Modes 4 and 5 take a third argument G and run the sharded path from ONE process on G ranks (gcb_multi_*, csrc/multi.cu; ranks
// are mapped round-robin onto the visible GPUs, so G > #GPUs still works: several slabs per device):
//   4 N G   BASELINE config 4 in miniature: the SVL lattice of config 3 on N^3, z-slabs, P2P min/max exchange
//   5 N G   BASELINE config 5 on several GPUs: STORED density + grid_points + colour field, owned layers only per rank, the +z halo
//           layer staged from the neighbour's memory by the extraction kernel
// and compare the concatenated rank meshes with the single-rank mesh byte for byte ("PARITY OK").
//
// `3 N --full [out.obj]` runs the WHOLE lattice workflow through the C ABI: unit cell -> spectrum -> period field -> phase solve ->
// field -> mesh -> .obj (Multitopo::unit_lattice + spatial_lattice_run).
//
//   gpucad_headless <config 1|2|3|5> [N] [out.obj]        gpucad_headless <4|5> N G        gpucad_headless 3 N --full [out.obj]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "gpucad_host.hpp"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

struct Mc {  // initMC_two
    uint3 gridSize, gridSizeShift, gridSizeMask;
    uint numVoxels, maxVerts;
    float3 voxelSize, gridcenter;
    uint *d_voxelVerts, *d_voxelVertsScan, *d_voxelOccupied, *d_voxelOccupiedScan, *d_compVoxelArray;
    float4 *d_pos, *d_normal;
    Mc(uint nx, uint ny, uint nz, float dx) {
        gridSize = make_uint3(nx, ny, nz);
        gridSizeMask = make_uint3(nx - 1, ny - 1, nz - 1);
        gridSizeShift = make_uint3(1, nx - 1, (nx - 1) * (ny - 1));
        numVoxels = gridSizeMask.x * gridSizeMask.y * gridSizeMask.z;
        voxelSize = make_float3(dx, dx, dx);
        gridcenter = make_float3(0, 0, 0);
        maxVerts = (uint)std::max<size_t>((size_t)nx * ny * nz * 4, 300000);  // main.cu:2850
        size_t m = sizeof(uint) * (size_t)numVoxels;
        CK(cudaMalloc(&d_voxelVerts, m)); CK(cudaMalloc(&d_voxelVertsScan, m)); CK(cudaMalloc(&d_voxelOccupied, m));
        CK(cudaMalloc(&d_voxelOccupiedScan, m)); CK(cudaMalloc(&d_compVoxelArray, m));
        CK(cudaMalloc(&d_pos, sizeof(float4) * (size_t)maxVerts)); CK(cudaMalloc(&d_normal, sizeof(float4) * (size_t)maxVerts));
    }
};

static float elapsed(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b); return ms; }

// ------------------------------------------------------------------ sharded modes (one process, G ranks)
static void mcheck(gcb_multi* m, int rc, const char* what) {
    if (rc != 0) { fprintf(stderr, "gpucad_b200 multi: %s failed: %s\n", what, gcb_multi_last_error(m)); exit(EXIT_FAILURE); }
}
static gcb_multi* make_multi(int G) {
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    std::vector<int> devs(G);
    for (int r = 0; r < G; ++r) devs[r] = r % ndev;
    gcb_multi* m = nullptr;
    if (gcb_multi_create(&m, G, devs.data()) != 0) { fprintf(stderr, "gcb_multi_create failed\n"); exit(EXIT_FAILURE); }
    return m;
}
struct HostMesh { std::vector<float> pos, norm; std::vector<unsigned> comp; unsigned long long active = 0, verts = 0; float ms = 0; };

// SVL lattice on G ranks; the same synthetic phases / coefficients as mode 3
static HostMesh svl_sharded(int N, int G) {
    const int R = 4, C = N / R, NH = 62;
    std::vector<float> phi((size_t)NH * C * C * C), coef(2 * NH);
    int h = 0;
    for (int k = -2; k <= 2 && h < NH; ++k) for (int j = -2; j <= 2 && h < NH; ++j) for (int i = -2; i <= 2 && h < NH; ++i, ++h) {
        coef[2 * h] = 0.05f * (1 + (h % 5)); coef[2 * h + 1] = 0.03f * ((h % 3) - 1);
        for (int z = 0; z < C; ++z) for (int y = 0; y < C; ++y) for (int x = 0; x < C; ++x)
            phi[(((size_t)h * C + z) * C + y) * C + x] = 6.2831853f / 10.0f * (i * (x - C / 2.0f) + j * (y - C / 2.0f) + k * (z - C / 2.0f)) + 0.01f * x * y / C;
    }
    gcb_multi* m = make_multi(G);
    std::vector<float*> d_svl(G), d_phi(G);
    std::vector<const float*> d_phi_c(G);
    std::vector<int> czl(G), cz0(G);
    std::vector<unsigned> z0(G), z1(G);
    for (int r = 0; r < G; ++r) {
        gcb_slab_bounds((unsigned)N, G, r, 2, &z0[r], &z1[r]);
        int c0, c1;
        gcb_control_slab(z0[r], z1[r], R, C, &c0, &c1);
        cz0[r] = c0; czl[r] = c1 - c0 + 1;
        CK(cudaSetDevice(gcb_multi_device(m, r)));
        const size_t per = (size_t)czl[r] * C * C;
        CK(cudaMalloc(&d_phi[r], (size_t)NH * per * 4));
        for (int hh = 0; hh < NH; ++hh)   // planes c0 .. c1 of every harmonic, back to back
            CK(cudaMemcpy(d_phi[r] + (size_t)hh * per, phi.data() + ((size_t)hh * C + c0) * C * C, per * 4, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&d_svl[r], (size_t)N * N * (z1[r] - z0[r] + 1) * 4));
        d_phi_c[r] = d_phi[r];
    }
    const gcb_float3 vs{0.25f, 0.25f, 0.25f}, gc{0, 0, 0};
    std::vector<unsigned long long> act(G), verts(G), off(G), cap(G);
    float mm[2];
    mcheck(m, gcb_multi_svl_lattice(m, d_svl.data(), d_phi_c.data(), NH, coef.data(), C, C, czl.data(), cz0.data(), N, N, (unsigned)N, 0.25f, 0.25f, 0.25f, 0.25f,
                                    0.20f, 0.30f, vs, gc, nullptr, nullptr, nullptr, 1, act.data(), verts.data(), off.data(), mm), "multi_svl_lattice(count)");
    std::vector<void*> pos(G), norm(G);
    for (int r = 0; r < G; ++r) {
        cap[r] = verts[r] + 3;   // count-then-allocate; +3: the `index < maxVerts - 3` guard
        CK(cudaSetDevice(gcb_multi_device(m, r)));
        CK(cudaMalloc(&pos[r], cap[r] * 16)); CK(cudaMalloc(&norm[r], cap[r] * 16));
    }
    HostMesh out;
    for (int rep = 0; rep < 3; ++rep) {
        mcheck(m, gcb_multi_svl_lattice(m, d_svl.data(), d_phi_c.data(), NH, coef.data(), C, C, czl.data(), cz0.data(), N, N, (unsigned)N, 0.25f, 0.25f, 0.25f, 0.25f,
                                        0.20f, 0.30f, vs, gc, pos.data(), norm.data(), cap.data(), 0, act.data(), verts.data(), off.data(), mm), "multi_svl_lattice");
        out.ms = gcb_multi_last_ms(m);
    }
    for (int r = 0; r < G; ++r) { out.active += act[r]; out.verts += verts[r]; }
    out.pos.resize(out.verts * 4); out.norm.resize(out.verts * 4);
    for (int r = 0; r < G; ++r) {
        CK(cudaSetDevice(gcb_multi_device(m, r)));
        CK(cudaMemcpy(out.pos.data() + off[r] * 4, pos[r], verts[r] * 16, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(out.norm.data() + off[r] * 4, norm[r], verts[r] * 16, cudaMemcpyDeviceToHost));
        cudaFree(pos[r]); cudaFree(norm[r]); cudaFree(d_svl[r]); cudaFree(d_phi[r]);
    }
    printf("  %d rank(s): activeVoxels=%llu totalVerts=%llu  field min/max %g %g  %.3f ms (max over ranks)\n", G, out.active, out.verts, mm[0], mm[1], out.ms);
    gcb_multi_destroy(m);
    return out;
}

// stored fields on G ranks: density + grid_points (stored crossing parameters, a solid patch) + colour field
static HostMesh density_sharded(int N, int G) {
    const int nx = N, ny = N / 2, nz = N / 2;
    const size_t layer = (size_t)nx * ny, n = layer * nz;
    std::vector<float> dens(n), result(n);
    std::vector<grid_points> gp(n);
    for (int z = 0; z < nz; ++z) for (int y = 0; y < ny; ++y) for (int x = 0; x < nx; ++x) {
        const size_t i = ((size_t)z * ny + y) * nx + x;
        const float v = 0.5f + 0.5f * sinf(0.105f * x) * sinf(0.085f * y + 0.4f) * cosf(0.095f * z);
        dens[i] = 0.07f + 0.93f * v * v;
        result[i] = 0.001f * (float)((x * 7 + y * 13 + z * 29) % 997);
        gp[i].val = ((x + 2 * y + 3 * z) % 41 == 0) ? -1 : 0;
        gp[i].t_x = ((x * 3 + z) % 5 == 0) ? 0.25f + 0.5f * (float)((y + z) % 3) / 3.0f : 0.0f;
        gp[i].t_y = ((y * 5 + x) % 7 == 0) ? 0.6f : 0.0f;
        gp[i].t_z = ((z * 11 + y) % 4 == 0) ? 0.1f + 0.2f * (float)(x % 4) : 0.0f;
    }
    gcb_multi* m = make_multi(G);
    std::vector<float*> d_dens(G), d_res(G);
    std::vector<gcb_grid_points*> d_gp(G);
    std::vector<unsigned*> d_comp(G);
    std::vector<void*> pos(G), norm(G);
    std::vector<unsigned long long> cap(G), act(G), verts(G), off(G), aoff(G);
    std::vector<unsigned> z0(G), z1(G);
    for (int r = 0; r < G; ++r) {
        gcb_slab_bounds((unsigned)nz, G, r, 2, &z0[r], &z1[r]);
        const size_t owned = (size_t)(z1[r] - z0[r] + (r == G - 1 ? 1 : 0)) * layer;   // OWNED layers only; the last rank owns its final layer too
        CK(cudaSetDevice(gcb_multi_device(m, r)));
        CK(cudaMalloc(&d_dens[r], owned * 4)); CK(cudaMalloc(&d_res[r], owned * 4)); CK(cudaMalloc(&d_gp[r], owned * sizeof(grid_points)));
        CK(cudaMemcpy(d_dens[r], dens.data() + z0[r] * layer, owned * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_res[r], result.data() + z0[r] * layer, owned * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_gp[r], gp.data() + z0[r] * layer, owned * sizeof(grid_points), cudaMemcpyHostToDevice));
        cap[r] = (unsigned long long)(z1[r] - z0[r]) * layer * 4 + 300000;
        CK(cudaMalloc(&pos[r], cap[r] * 16)); CK(cudaMalloc(&norm[r], cap[r] * 16));
        CK(cudaMalloc(&d_comp[r], (size_t)(z1[r] - z0[r]) * (nx - 1) * (ny - 1) * 4));
    }
    const gcb_uint3 gs{(unsigned)nx, (unsigned)ny, (unsigned)nz};
    const gcb_float3 vs{0.5f, 0.5f, 0.5f}, gc{(float)nx / 4, 1.5f, (float)nz / 4 + 0.5f};
    HostMesh out;
    for (int rep = 0; rep < 3; ++rep) {
        mcheck(m, gcb_multi_computeIsosurface_2(m, d_gp.data(), d_dens.data(), d_res.data(), gs, vs, gc, 0.4f, 0.0f, pos.data(), norm.data(), cap.data(), d_comp.data(),
                                                act.data(), verts.data(), off.data(), aoff.data()), "multi_computeIsosurface_2");
        out.ms = gcb_multi_last_ms(m);
    }
    for (int r = 0; r < G; ++r) { out.active += act[r]; out.verts += verts[r]; }
    out.pos.resize(out.verts * 4); out.norm.resize(out.verts * 4); out.comp.resize(out.active);
    for (int r = 0; r < G; ++r) {
        CK(cudaSetDevice(gcb_multi_device(m, r)));
        CK(cudaMemcpy(out.pos.data() + off[r] * 4, pos[r], verts[r] * 16, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(out.norm.data() + off[r] * 4, norm[r], verts[r] * 16, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(out.comp.data() + aoff[r], d_comp[r], act[r] * 4, cudaMemcpyDeviceToHost));
        cudaFree(pos[r]); cudaFree(norm[r]); cudaFree(d_comp[r]); cudaFree(d_dens[r]); cudaFree(d_res[r]); cudaFree(d_gp[r]);
    }
    printf("  %d rank(s): activeVoxels=%llu totalVerts=%llu  %.3f ms (max over ranks)\n", G, out.active, out.verts, out.ms);
    gcb_multi_destroy(m);
    return out;
}

// ------------------------------------------------------------------ mode 3 --full: the whole SVL workflow through the C ABI
// Multitopo::unit_lattice (main.cu:3577-3706) + spatial_lattice_run (:3904-4037): unit cell -> spectrum -> period field -> normalise_three
// -> phase solve of all 62 harmonics -> fused field + extraction on the 2x refined grid (the app's own ratio) -> .obj.
static int full_workflow(int N, const char* obj) {
    const int NU = 61, NH = 62, C = N / 2;   // unit cell size of the reference (main.cu:582), control grid = coarse grid of the app
    const size_t nc = (size_t)C * C * C, nf = (size_t)N * N * N;
    gcb_ctx* ctx = gpucad::ctx();
    cudaEvent_t ev[8];
    for (auto& e : ev) cudaEventCreate(&e);
    float *d_cell, *d_spec, *d_period, *d_phi, *d_svl;
    CK(cudaMalloc(&d_cell, (size_t)NU * NU * NU * 4)); CK(cudaMalloc(&d_spec, 125 * 8)); CK(cudaMalloc(&d_period, nc * 4));
    CK(cudaMalloc(&d_phi, (size_t)NH * nc * 4)); CK(cudaMalloc(&d_svl, nf * 4));
    std::vector<int> ijk;
    for (int k = -2; k <= 2; ++k) for (int j = -2; j <= 2; ++j) for (int i = -2; i <= 2; ++i) if ((int)ijk.size() < 3 * NH) { ijk.push_back(i); ijk.push_back(j); ijk.push_back(k); }
    cudaEventRecord(ev[0]);
    gpucad::check(gcb_create_lattice(ctx, d_cell, NU, NU, NU, NU * NU * NU, 0), "create_lattice");
    gpucad::check(gcb_unit_lattice_spectrum(ctx, d_cell, NU, NU, NU, 2, d_spec), "unit_lattice_spectrum");
    std::vector<float> spec(250);
    CK(cudaMemcpy(spec.data(), d_spec, 250 * 4, cudaMemcpyDeviceToHost));
    cudaEventRecord(ev[1]);
    Gratings latt;  // the reference's own two calls (main.cu:3927-3931), through the host mirror
    latt.period_data(d_period, C, C, C, 1.f, 1.f, 1.f, C / 2.0f, C / 2.0f, C / 2.0f, 'z');
    latt.GPU_buffer_normalise_three(d_period, d_period, nc, (float)(C / 10), (float)(C / 4));
    cudaEventRecord(ev[2]);
    std::vector<int> fi(NH);
    std::vector<float> fr(NH);
    gpucad::check(gcb_svl_phase_solve(ctx, d_phi, d_period, NH, ijk.data(), C, C, C, 1.f, 1.f, 1.f, 'r', 2, 8.f, 8.f, 8.f, 8.f, 0.5f, 0.05f, 0, 500, 0.01f, fi.data(), fr.data()),
                  "svl_phase_solve");
    cudaEventRecord(ev[3]);
    unsigned long long a64 = 0, t64 = 0;
    float mm[2];
    const gcb_float3 vs{0.5f, 0.5f, 0.5f}, gc{0, 0, 0};
    gpucad::check(gcb_svl_lattice(ctx, d_svl, d_phi, NH, spec.data(), C, C, C, N, N, N, 0.5f, 0.5f, 0.5f, 0.25f, 0.20f, 0.30f, vs, gc, nullptr, nullptr, 3, &a64, &t64, mm),
                  "svl_lattice(count)");
    float4 *pos, *norm;
    CK(cudaMalloc(&pos, (t64 + 3) * 16)); CK(cudaMalloc(&norm, (t64 + 3) * 16));
    cudaEventRecord(ev[4]);
    gpucad::check(gcb_svl_lattice(ctx, d_svl, d_phi, NH, spec.data(), C, C, C, N, N, N, 0.5f, 0.5f, 0.5f, 0.25f, 0.20f, 0.30f, vs, gc, pos, norm, t64 + 3, &a64, &t64, mm),
                  "svl_lattice");
    cudaEventRecord(ev[5]);
    long iters = 0;
    for (int h = 0; h < NH; ++h) iters += fi[h];
    printf("config 3 --full N=%d (control %d^3): unit cell + spectrum %.3f ms, period field %.3f ms, phase solve %.3f ms (%ld CG iterations over %d harmonics), "
           "field + extraction %.3f ms\n", N, C, elapsed(ev[0], ev[1]), elapsed(ev[1], ev[2]), elapsed(ev[2], ev[3]), iters, NH, elapsed(ev[4], ev[5]));
    printf("config 3 N=%d: activeVoxels=%llu totalVerts=%llu triangles=%llu  %.3f ms  (field min/max %g %g)\n", N, a64, t64, t64 / 3,
           elapsed(ev[0], ev[3]) + elapsed(ev[4], ev[5]), mm[0], mm[1]);
    if (obj && t64) {
        gpucad::check(gcb_file_write_obj(ctx, pos, (unsigned)t64, obj), "file_write_obj");
        printf("wrote %s\n", obj);
    }
    return 0;
}

static int sharded_mode(int config, int N, int G) {
    if (G < 1 || G > 16) { fprintf(stderr, "G must be 1..16\n"); return 2; }
    printf("config %d sharded: N=%d, %d rank(s) vs 1 rank\n", config, N, G);
    const HostMesh one = config == 4 ? svl_sharded(N, 1) : density_sharded(N, 1);
    const HostMesh many = config == 4 ? svl_sharded(N, G) : density_sharded(N, G);
    const bool ok = one.active == many.active && one.verts == many.verts && one.verts > 0 &&
                    memcmp(one.pos.data(), many.pos.data(), one.pos.size() * 4) == 0 && memcmp(one.norm.data(), many.norm.data(), one.norm.size() * 4) == 0 &&
                    one.comp == many.comp;
    printf("%s: concatenated meshes of %d ranks %s the single-rank mesh (%llu vertices, %zu bytes compared); 1 rank %.3f ms, %d ranks %.3f ms\n",
           ok ? "PARITY OK" : "PARITY FAIL", G, ok ? "equal" : "DIFFER from", one.verts, one.pos.size() * 8, one.ms, G, many.ms);
    return ok ? 0 : 1;
}

int main(int argc, char** argv) {
    const int config = argc > 1 ? atoi(argv[1]) : 1;
    const int N = argc > 2 ? atoi(argv[2]) : (config == 1 ? 128 : 256);
    if (config == 4 || (config == 5 && argc > 3 && atoi(argv[3]) > 0 && strchr(argv[3], '.') == nullptr))
        return sharded_mode(config, N, argc > 3 ? atoi(argv[3]) : 2);
    const char* obj = argc > 3 ? argv[3] : nullptr;
    if (config == 3)
        for (int i = 2; i < argc; ++i)
            if (!strcmp(argv[i], "--full")) {
                const char* o = nullptr;
                for (int j = 3; j < argc; ++j) if (j != i) o = argv[j];
                return full_workflow(N, o);
            }
    Isosurface isosurf;
    Gratings lattice;
    Fft_lattice fftlattice;
    File_output output_file;
    uint *d_tri, *d_nv;
    isosurf.allocateTextures_s(&d_tri, &d_nv);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    uint active = 0, total = 0;
    float ms = 0;
    float4* d_pos = nullptr;

    if (config == 1) {  // gyroid unit cell, band [0.20, 0.30], mask iso 0.25
        Mc mc(N, N, N, 1.0f);
        size_t n = (size_t)N * N * N;
        float *d_vol, *d_mask;
        CK(cudaMalloc(&d_vol, n * 4)); CK(cudaMalloc(&d_mask, n * 4));
        float iso = 0.25f;
        cudaEventRecord(e0);
        fftlattice.create_lattice(d_vol, N, N, N, (uint)n, 0);                                         // unit_latticeone :3717
        lattice.GPU_buffer_normalise_buffer(d_vol, d_vol, (int)n);                                     // :3719
        lattice.GPU_buffer_normalise_four(d_vol, d_mask, d_vol, n, N, N, N, 0.20f, 0.30f);            // display_unit_lattice :4113
        isosurf.computeIsosurface_latticeone(d_mask, mc.d_pos, mc.d_normal, iso, mc.numVoxels, mc.d_voxelVerts, mc.d_voxelVertsScan, mc.d_voxelOccupied,
                                             mc.d_voxelOccupiedScan, mc.gridSize, mc.gridSizeShift, mc.gridSizeMask, mc.voxelSize, mc.gridcenter, &active,
                                             &total, mc.d_compVoxelArray, mc.maxVerts, d_vol, 0.20f, 0.30f);     // :4115-4119
        cudaEventRecord(e1);
        ms = elapsed(e0, e1);
        d_pos = mc.d_pos;
    } else if (config == 2) {  // sphere U box - cylinder on the 2x refined grid (dx2 = 0.5)
        Mc mc(N, N, N, 0.5f);
        size_t n = (size_t)N * N * N;
        const float s = N / 256.0f;
        Modelling model(N, N, N);
        float *d_boundary, *d_lat;
        grid_points* vol_one;
        CK(cudaMalloc(&d_boundary, n * 4)); CK(cudaMalloc(&d_lat, n * 4)); CK(cudaMalloc(&vol_one, n * sizeof(grid_points)));
        CK(cudaMemset(d_boundary, 0, n * 4)); CK(cudaMemset(d_lat, 0, n * 4)); CK(cudaMemset(vol_one, 0, n * sizeof(grid_points)));  // init_Boundary :3282
        size_t nfacets = 0;
        cudaEventRecord(e0);
        model.sphere_with_center(d_boundary, make_float3(0, 0, 0), 40 * s, 2, N, N, N, 0.5f, 0.5f, 0.5f, false);
        isosurf.copy_parameter(mc.d_voxelVerts, 0.0f, mc.gridSize, mc.gridSizeShift, mc.gridSizeMask, mc.voxelSize, mc.numVoxels, vol_one, d_boundary, d_lat,
                               false, false, 0.20f, 0.30f, true, false, false);                                    // retain :3309
        model.cuboid(d_boundary, make_float3(0, 0, 0), make_float3(0.3f, 0.2f, 0.1f), 90 * s, 50 * s, 60 * s, N, N, N, 0.5f, 0.5f, 0.5f);
        isosurf.copy_parameter(mc.d_voxelVerts, 0.0f, mc.gridSize, mc.gridSizeShift, mc.gridSizeMask, mc.voxelSize, mc.numVoxels, vol_one, d_boundary, d_lat,
                               false, false, 0.20f, 0.30f, true, false, false);
        model.distance_from_line(d_boundary, make_float3(0, 0, 0), make_float3(0, 0, 1), 18 * s, 2, 200 * s, N, N, N, 0.5f, 0.5f, 0.5f, false);
        isosurf.computeIsosurface(nullptr, mc.gridSize, mc.d_pos, mc.d_normal, 0.0f, mc.numVoxels, mc.d_voxelVerts, mc.d_voxelVertsScan, mc.d_voxelOccupied,
                                  mc.d_voxelOccupiedScan, mc.gridSize, mc.gridSizeShift, mc.gridSizeMask, mc.voxelSize, mc.gridcenter, &active, &total,
                                  mc.d_compVoxelArray, mc.maxVerts, vol_one, d_boundary, nullptr, d_lat, 0.20f, 0.30f, false, true, false, true, false, false,
                                  false, false, false, &nfacets);                                                  // show_model :3410-3413, obj_diff
        cudaEventRecord(e1);
        ms = elapsed(e0, e1);
        d_pos = mc.d_pos;
    } else if (config == 3) {  // fused SVL lattice: control grid N/4, 62 harmonics, synthetic linear phases
        const int C = N / 4, NH = 62;
        std::vector<float> phi((size_t)NH * C * C * C), coef(2 * NH);
        int h = 0;
        for (int k = -2; k <= 2 && h < NH; ++k) for (int j = -2; j <= 2 && h < NH; ++j) for (int i = -2; i <= 2 && h < NH; ++i, ++h) {
            coef[2 * h] = 0.05f * (1 + (h % 5)); coef[2 * h + 1] = 0.03f * ((h % 3) - 1);
            for (int z = 0; z < C; ++z) for (int y = 0; y < C; ++y) for (int x = 0; x < C; ++x)
                phi[(((size_t)h * C + z) * C + y) * C + x] = 6.2831853f / 10.0f * (i * (x - C / 2.0f) + j * (y - C / 2.0f) + k * (z - C / 2.0f));
        }
        float *d_phi, *d_svl, *h_phi;
        size_t n = (size_t)N * N * N;
        CK(cudaMalloc(&d_phi, phi.size() * 4)); CK(cudaMalloc(&d_svl, n * 4)); CK(cudaMallocHost(&h_phi, phi.size() * 4));
        memcpy(h_phi, phi.data(), phi.size() * 4);
        unsigned long long a64 = 0, t64 = 0;
        float mm[2];
        gcb_float3 vs{0.25f, 0.25f, 0.25f}, gc{0, 0, 0};
        // count first (mesh stays unallocated), then allocate exactly: SURVEY.md 7 "Capacity"
        gpucad::check(gcb_svl_field(gpucad::ctx(), d_svl, d_phi, 0, coef.data(), C, C, C, 0, N, N, N, gcb_slab{0, (unsigned)N}, 0.25f, 0.25f, 0.25f, 0, nullptr), "warm");
        CK(cudaMemcpy(d_phi, h_phi, phi.size() * 4, cudaMemcpyHostToDevice));
        float4 *pos = nullptr, *norm = nullptr;
        gpucad::check(gcb_svl_lattice(gpucad::ctx(), d_svl, d_phi, NH, coef.data(), C, C, C, N, N, N, 0.25f, 0.25f, 0.25f, 0.25f, 0.20f, 0.30f, vs, gc, nullptr,
                                      nullptr, 3, &a64, &t64, mm), "svl_lattice(count)");
        CK(cudaMalloc(&pos, (t64 + 3) * 16)); CK(cudaMalloc(&norm, (t64 + 3) * 16));
        cudaEventRecord(e0);
        gpucad::check(gcb_svl_lattice_host(gpucad::ctx(), h_phi, d_phi, d_svl, NH, coef.data(), C, C, C, N, N, N, 0.25f, 0.25f, 0.25f, 0.25f, 0.20f, 0.30f, vs, gc,
                                           pos, norm, t64 + 3, &a64, &t64, mm), "svl_lattice_host");
        cudaEventRecord(e1);
        ms = elapsed(e0, e1);
        active = (uint)a64; total = (uint)t64; d_pos = pos;
        printf("field min/max %g %g\n", mm[0], mm[1]);
    } else if (config == 5) {  // density (coarse N/2 x N/4 x N/4 blobs) -> refine -> computeIsosurface_2, iso 0.4
        const int nx = N, ny = N / 2, nz = N / 2, cx = nx / 2, cy = ny / 2, cz = nz / 2;
        Mc mc(nx, ny, nz, 0.5f);
        std::vector<float> coarse((size_t)cx * cy * cz);
        for (int z = 0; z < cz; ++z) for (int y = 0; y < cy; ++y) for (int x = 0; x < cx; ++x) {
            float v = 0.5f + 0.5f * sinf(0.21f * x) * sinf(0.17f * y + 0.4f) * cosf(0.19f * z);
            coarse[((size_t)z * cy + y) * cx + x] = 0.07f + 0.93f * v * v;
        }
        float *d_coarse, *d_pitched, *d_dens, *d_result;
        grid_points* vol_topo;
        size_t n = (size_t)nx * ny * nz;
        CK(cudaMalloc(&d_coarse, coarse.size() * 4)); CK(cudaMalloc(&d_pitched, coarse.size() * 4)); CK(cudaMalloc(&d_dens, n * 4));
        CK(cudaMalloc(&d_result, n * 4)); CK(cudaMalloc(&vol_topo, n * sizeof(grid_points)));
        CK(cudaMemcpy(d_coarse, coarse.data(), coarse.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemset(d_result, 0, n * 4)); CK(cudaMemset(vol_topo, 0, n * sizeof(grid_points)));
        lattice.setupTexture(cx, cy, cz);
        cudaPitchedPtr pp = make_cudaPitchedPtr(d_pitched, cx * 4, cx * 4, cy);
        cudaEventRecord(e0);
        lattice.copytotexture(d_coarse, pp, cx, cy, cz);                                               // toprun_struct :3060
        lattice.updateTexture(pp);
        lattice.refine(d_dens, nx, ny, nz, 0.5f, 0.5f, 0.5f);                                          // :3064
        isosurf.patch_topo_field(d_dens, nx, ny, nz, vol_topo);                                        // :3066
        isosurf.computeIsosurface_2(mc.d_pos, mc.d_normal, 0.4f, mc.numVoxels, mc.d_voxelVerts, mc.d_voxelVertsScan, mc.d_voxelOccupied, mc.d_voxelOccupiedScan,
                                    mc.gridSize, mc.gridSizeShift, mc.gridSizeMask, mc.voxelSize, mc.gridcenter, &active, &total, mc.d_compVoxelArray, mc.maxVerts,
                                    vol_topo, vol_topo, d_dens, d_dens, 0.0f, d_result, nullptr);      // :3107-3109
        cudaEventRecord(e1);
        ms = elapsed(e0, e1);
        d_pos = mc.d_pos;
    } else {
        fprintf(stderr, "usage: gpucad_headless <1|2|3|5> [N] [out.obj]   |   gpucad_headless <4|5> N G\n");
        return 2;
    }
    printf("config %d N=%d: activeVoxels=%u totalVerts=%u triangles=%u  %.3f ms  (%.1f Mvoxel/s)\n", config, N, active, total, total / 3, ms,
           (double)N * N * N / ms / 1e3);
    if (obj && total) {
        output_file.file_write_obj(d_pos, total, obj);
        printf("wrote %s\n", obj);
    }
    return 0;
}
