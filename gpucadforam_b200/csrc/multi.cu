// multi.cu -- z-slab sharding over several GPUs from ONE host process (C ABI: gcb_multi_*, declared in include/gpucad_b200.h).
//
// The reference is single-GPU (no multi-device code at all); the contract here is BASELINE.json's: the grid is cut
// into z-slabs, every rank evaluates / holds the point layers of its cells plus the +z halo layer, the global min/max joins the
// ranks between field and extraction, and per-rank counts are exclusive-scanned so that the rank meshes, concatenated in rank
// order, are the single-GPU mesh byte for byte (vertex order = ascending linear cell id, z slowest; MarchingCubes_kernel.cu:120-136,
// :2160).  One context and one stream per rank; everything is enqueued without host synchronisation until the counts are read:
//   * analytic fields (configs 3 / 4): each rank evaluates its halo layer itself; the 2-float min/max exchange is a one-warp
//     kernel per rank that reads every rank's pair through peer-mapped pointers (NVLink P2P) after cross-device event waits;
//   * stored fields (config 5 on several GPUs): a rank holds its OWNED layers only and the extraction kernel stages the halo
//     layer straight from the upper neighbour's buffer (second TMA bulk copy of the tile, mc_extract.cu) -- no exchange step.
// bench.py's N > 1 runs use one process per GPU with torch.distributed / NCCL for the same two exchanges (sharding.py); this file
// is the C++ host of the same scheme for callers that own all GPUs of a box in one process (host/headless_main.cpp modes 4, 5).
#include "common.cuh"

#include <cmath>
#include <vector>

namespace gcb {

constexpr int kMaxRanks = 16;
struct PeerPairs { const float* p[kMaxRanks]; };

// out = {min_j pair_j[0], max_j pair_j[1]}: every rank's {min, max} read where it lies (own memory or a peer's over NVLink)
__global__ void combine_minmax_kernel(const PeerPairs pp, int n, float* __restrict__ out) {
    if (threadIdx.x == 0) {
        float lo = pp.p[0][0], hi = pp.p[0][1];
        for (int j = 1; j < n; ++j) {
            lo = fminf(lo, pp.p[j][0]);
            hi = fmaxf(hi, pp.p[j][1]);
        }
        out[0] = lo;
        out[1] = hi;
    }
}

struct Multi {
    int n = 0;
    std::vector<int> dev;
    std::vector<gcb_ctx*> ctx;
    std::vector<cudaStream_t> stream;
    std::vector<cudaEvent_t> ev_ready, ev_field, ev_t0, ev_t1;
    std::vector<float*> d_mm_local;   // per rank: {min, max} of its slab (decoded floats)
    std::vector<float*> d_ab;         // per rank: global {min, max}
    std::vector<unsigned long long*> h_totals;  // per rank, pinned: {active, vertices}
    float* h_mm = nullptr;                      // pinned: global {min, max} for the caller
    std::string err;
    float last_ms = -1.f;
};

static int mfail(Multi* m, const std::string& s) { m->err = s; return 1; }
struct DeviceGuard {  // a sharded call hops between devices: leave the caller's current device as it was, on every exit path
    int prev = -1;
    DeviceGuard() { cudaGetDevice(&prev); }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define MCK(m, call)                                                                                         \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess) return mfail((m), std::string(#call) + ": " + cudaGetErrorString(e_));        \
    } while (0)

// cell layers [z0, z1) of rank `rank`: interior cuts are multiples of `align`, nearest to the even split (ties to even, like
// Python's round() in gpucadforam_b200/sharding.py slab_bounds -- the two must agree, tests/test_sharding_gloo.py checks it)
static void slab_cut(unsigned gnz, int world, int rank, unsigned align, unsigned* z0, unsigned* z1) {
    const unsigned cells = gnz ? gnz - 1 : 0;
    auto cut = [&](int r) -> unsigned {
        if (r <= 0) return 0u;
        if (r >= world) return cells;
        const double q = std::nearbyint((double)r * (double)cells / (double)world / (double)align) * (double)align;
        return (unsigned)std::min<double>((double)cells, std::max(0.0, q));
    };
    *z0 = cut(rank);
    *z1 = cut(rank + 1);
}

}  // namespace gcb

using namespace gcb;

extern "C" {

int gcb_slab_bounds(unsigned int gnz, int world, int rank, unsigned int align, unsigned int* z0, unsigned int* z1) {
    if (!z0 || !z1 || world < 1 || rank < 0 || rank >= world || align < 1) return 1;
    slab_cut(gnz, world, rank, align, z0, z1);
    return 0;
}
int gcb_control_slab(unsigned int z0, unsigned int z1, int ratio, int cz_global, int* c0, int* c1) {
    if (!c0 || !c1 || ratio < 1 || cz_global < 1) return 1;
    *c0 = (int)(z0 / (unsigned)ratio);
    *c1 = std::min((int)(z1 / (unsigned)ratio) + 1, cz_global - 1);
    return 0;
}

int gcb_multi_create(gcb_multi** out, int n, const int* devices) {
    if (!out || n < 1 || n > kMaxRanks || !devices) return 1;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return 2;  // no CPU fallback
    Multi* m = new (std::nothrow) Multi();
    if (!m) return 1;
    m->n = n;
    int prev = 0;
    cudaGetDevice(&prev);
    bool ok = true;
    for (int r = 0; r < n && ok; ++r) {
        const int d = devices[r];
        if (d < 0 || d >= ndev) { ok = false; break; }
        ok = cudaSetDevice(d) == cudaSuccess;
        cudaStream_t st = nullptr;
        ok = ok && cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
        gcb_ctx* c = nullptr;
        ok = ok && gcb_create(&c, d, st) == 0;
        cudaEvent_t e[4] = {nullptr, nullptr, nullptr, nullptr};
        for (int i = 0; i < 2 && ok; ++i) ok = cudaEventCreateWithFlags(&e[i], cudaEventDisableTiming) == cudaSuccess;
        for (int i = 2; i < 4 && ok; ++i) ok = cudaEventCreate(&e[i]) == cudaSuccess;
        float *mm = nullptr, *ab = nullptr;
        unsigned long long* ht = nullptr;
        ok = ok && cudaMalloc(&mm, 16) == cudaSuccess && cudaMalloc(&ab, 16) == cudaSuccess && cudaMallocHost(&ht, 16) == cudaSuccess;
        m->dev.push_back(d); m->ctx.push_back(c); m->stream.push_back(st);
        m->ev_ready.push_back(e[0]); m->ev_field.push_back(e[1]); m->ev_t0.push_back(e[2]); m->ev_t1.push_back(e[3]);
        m->d_mm_local.push_back(mm); m->d_ab.push_back(ab); m->h_totals.push_back(ht);
    }
    // peer access between every pair of DISTINCT devices (ranks may share a device: slabs one after another, e.g. for tests)
    for (int a = 0; a < n && ok; ++a)
        for (int b = 0; b < n && ok; ++b) {
            if (m->dev[a] == m->dev[b]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, m->dev[a], m->dev[b]);
            if (!can) { m->err = "no peer access between the devices"; ok = false; break; }
            cudaSetDevice(m->dev[a]);
            const cudaError_t e = cudaDeviceEnablePeerAccess(m->dev[b], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = false;
            cudaGetLastError();
        }
    cudaSetDevice(prev);
    if (!ok) { gcb_multi_destroy(reinterpret_cast<gcb_multi*>(m)); return 3; }
    *out = reinterpret_cast<gcb_multi*>(m);
    return 0;
}
int gcb_multi_destroy(gcb_multi* mm) {
    Multi* m = reinterpret_cast<Multi*>(mm);
    if (!m) return 1;
    int prev = 0;
    cudaGetDevice(&prev);
    for (size_t r = 0; r < m->ctx.size(); ++r) {
        cudaSetDevice(m->dev[r]);
        if (m->ctx[r]) gcb_destroy(m->ctx[r]);
        if (m->stream[r]) cudaStreamDestroy(m->stream[r]);
        for (cudaEvent_t e : {m->ev_ready[r], m->ev_field[r], m->ev_t0[r], m->ev_t1[r]}) if (e) cudaEventDestroy(e);
        cudaFree(m->d_mm_local[r]); cudaFree(m->d_ab[r]);
        if (m->h_totals[r]) cudaFreeHost(m->h_totals[r]);
    }
    if (m->h_mm) cudaFreeHost(m->h_mm);
    cudaSetDevice(prev);
    delete m;
    return 0;
}
int gcb_multi_size(gcb_multi* mm) { Multi* m = reinterpret_cast<Multi*>(mm); return m ? m->n : 0; }
gcb_ctx* gcb_multi_ctx(gcb_multi* mm, int rank) { Multi* m = reinterpret_cast<Multi*>(mm); return (m && rank >= 0 && rank < m->n) ? m->ctx[rank] : nullptr; }
int gcb_multi_device(gcb_multi* mm, int rank) { Multi* m = reinterpret_cast<Multi*>(mm); return (m && rank >= 0 && rank < m->n) ? m->dev[rank] : -1; }
const char* gcb_multi_last_error(gcb_multi* mm) { Multi* m = reinterpret_cast<Multi*>(mm); return m ? m->err.c_str() : "null handle"; }
float gcb_multi_last_ms(gcb_multi* mm) { Multi* m = reinterpret_cast<Multi*>(mm); return m ? m->last_ms : -1.f; }

// end of a sharded call: wait for every rank, read the counts, exclusive-scan them into global offsets, device time = max over ranks
static int finish(Multi* m, unsigned long long* active, unsigned long long* verts, unsigned long long* vert_offsets, unsigned long long* active_offsets) {
    float worst = 0.f;
    for (int r = 0; r < m->n; ++r) {
        MCK(m, cudaSetDevice(m->dev[r]));
        MCK(m, cudaStreamSynchronize(m->stream[r]));
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, m->ev_t0[r], m->ev_t1[r]) == cudaSuccess) worst = std::max(worst, ms);
    }
    m->last_ms = worst;
    unsigned long long vo = 0, ao = 0;
    for (int r = 0; r < m->n; ++r) {
        const unsigned long long a = m->h_totals[r][0], v = a ? m->h_totals[r][1] : 0ull;
        if (active) active[r] = a;
        if (verts) verts[r] = v;
        if (vert_offsets) vert_offsets[r] = vo;
        if (active_offsets) active_offsets[r] = ao;
        vo += v; ao += a;
    }
    return 0;
}

int gcb_multi_svl_lattice(gcb_multi* mm, float* const* d_svl, const float* const* d_phi, int nh, const float* coef_host, int cx, int cy, const int* cz_local,
                          const int* cz0, int NX2, int NY2, unsigned int gnz, float dx, float dy, float dz, float isoValue, float isovalue1, float isovalue2,
                          gcb_float3 voxelSize, gcb_float3 gridcenter, void* const* pos, void* const* norm, const unsigned long long* max_verts, int count_only,
                          unsigned long long* active, unsigned long long* verts, unsigned long long* vert_offsets, float* minmax_out) {
    Multi* m = reinterpret_cast<Multi*>(mm);
    if (!m) return 1;
    if (!d_svl || !d_phi || !cz_local || !cz0 || (!count_only && (!pos || !norm || !max_verts))) return mfail(m, "multi_svl_lattice: null argument array");
    DeviceGuard guard;
    std::vector<unsigned> z0(m->n), z1(m->n);
    for (int r = 0; r < m->n; ++r) {
        slab_cut(gnz, m->n, r, 2u, &z0[r], &z1[r]);
        if (z1[r] <= z0[r]) return mfail(m, "multi_svl_lattice: more ranks than aligned cell layers");
    }
    // 1. field of every slab (its +z halo layer included: analytic field, recomputed, nothing exchanged) + slab min/max
    for (int r = 0; r < m->n; ++r) {
        MCK(m, cudaSetDevice(m->dev[r]));
        MCK(m, cudaEventRecord(m->ev_t0[r], m->stream[r]));
        const int nzl = (int)(z1[r] - z0[r] + 1);
        if (gcb_svl_field(m->ctx[r], d_svl[r], d_phi[r], nh, coef_host, cx, cy, cz_local[r], cz0[r], NX2, NY2, nzl, gcb_slab{z0[r], gnz}, dx, dy, dz, 0, m->d_mm_local[r]))
            return mfail(m, std::string("multi_svl_lattice rank ") + std::to_string(r) + ": " + gcb_last_error(m->ctx[r]));
        MCK(m, cudaEventRecord(m->ev_field[r], m->stream[r]));
    }
    // 2. global min/max on every rank: wait for the other ranks' fields (cross-device events), then read their pairs in place
    PeerPairs pp{};
    for (int j = 0; j < m->n; ++j) pp.p[j] = m->d_mm_local[j];
    for (int r = 0; r < m->n; ++r) {
        MCK(m, cudaSetDevice(m->dev[r]));
        for (int j = 0; j < m->n; ++j)
            if (j != r) MCK(m, cudaStreamWaitEvent(m->stream[r], m->ev_field[j], 0));
        combine_minmax_kernel<<<1, 32, 0, m->stream[r]>>>(pp, m->n, m->d_ab[r]);
        MCK(m, cudaGetLastError());
    }
    // 3. extraction of every slab with the global range read from device memory, global z / domain faces, counts to pinned memory
    for (int r = 0; r < m->n; ++r) {
        MCK(m, cudaSetDevice(m->dev[r]));
        Ctx* C = reinterpret_cast<Ctx*>(m->ctx[r]);
        C->launches++;  // the combine kernel above
        const gcb_uint3 gs{(unsigned)NX2, (unsigned)NY2, z1[r] - z0[r] + 1};
        if (gcb_internal_extract_band_raw(C, d_svl[r], 0.f, 0.f, m->d_ab[r], isoValue, isovalue1, isovalue2, gs, gcb_slab{z0[r], gnz}, voxelSize, gridcenter,
                                          count_only ? nullptr : pos[r], count_only ? nullptr : norm[r], count_only ? 0ull : max_verts[r], nullptr, count_only, nullptr,
                                          nullptr, m->h_totals[r]))
            return mfail(m, std::string("multi_svl_lattice rank ") + std::to_string(r) + ": " + gcb_last_error(m->ctx[r]));
        MCK(m, cudaEventRecord(m->ev_t1[r], m->stream[r]));
    }
    if (minmax_out) {
        MCK(m, cudaSetDevice(m->dev[0]));
        if (!m->h_mm) MCK(m, cudaMallocHost(&m->h_mm, 16));
        MCK(m, cudaMemcpyAsync(m->h_mm, m->d_ab[0], 2 * sizeof(float), cudaMemcpyDeviceToHost, m->stream[0]));
    }
    const int rc = finish(m, active, verts, vert_offsets, nullptr);
    if (rc) return rc;
    if (count_only && verts)  // count_only reports the vertex count even when it is zero-active (same as gcb_extract_band_raw)
        for (int r = 0; r < m->n; ++r) verts[r] = m->h_totals[r][1];
    if (minmax_out) { minmax_out[0] = m->h_mm[0]; minmax_out[1] = m->h_mm[1]; }
    return 0;
}

int gcb_multi_computeIsosurface_2(gcb_multi* mm, gcb_grid_points* const* vol_topo, float* const* vol_two, float* const* d_result, gcb_uint3 gridSizeGlobal,
                                  gcb_float3 voxelSize, gcb_float3 gridcenter, float isoValue, float isovalue1, void* const* pos, void* const* norm,
                                  const unsigned long long* max_verts, unsigned int* const* d_compVoxelArray, unsigned long long* active,
                                  unsigned long long* verts, unsigned long long* vert_offsets, unsigned long long* active_offsets) {
    Multi* m = reinterpret_cast<Multi*>(mm);
    if (!m) return 1;
    if (!vol_two || !pos || !norm || !max_verts) return mfail(m, "multi_computeIsosurface_2: null argument array");
    DeviceGuard guard;
    const unsigned gnz = gridSizeGlobal.z;
    const size_t layer = (size_t)gridSizeGlobal.x * gridSizeGlobal.y;
    std::vector<unsigned> z0(m->n), z1(m->n);
    for (int r = 0; r < m->n; ++r) {
        slab_cut(gnz, m->n, r, 2u, &z0[r], &z1[r]);
        if (z1[r] <= z0[r]) return mfail(m, "multi_computeIsosurface_2: more ranks than aligned cell layers");
    }
    // whatever each rank queued on its stream before this call (upload, refine, retain ...) produced its layers: mark that point
    for (int r = 0; r < m->n; ++r) {
        MCK(m, cudaSetDevice(m->dev[r]));
        MCK(m, cudaEventRecord(m->ev_ready[r], m->stream[r]));
    }
    for (int r = 0; r < m->n; ++r) {
        MCK(m, cudaSetDevice(m->dev[r]));
        Ctx* C = reinterpret_cast<Ctx*>(m->ctx[r]);
        MCK(m, cudaEventRecord(m->ev_t0[r], m->stream[r]));
        McArgs a;
        const gcb_uint3 gs{gridSizeGlobal.x, gridSizeGlobal.y, z1[r] - z0[r] + 1};
        base_args(a, M_TOPO, gs, voxelSize, gridcenter, isoValue);  // gz0 below makes vertex z and cell ids global: (float)(z + z0) - center.z
        a.iso1 = isovalue1;
        a.f0 = vol_two[r];
        a.f1 = d_result ? d_result[r] : nullptr;
        a.gp = vol_topo ? (const GridPoint*)vol_topo[r] : nullptr;
        if (r + 1 < m->n) {
            // the halo layer z1 is the first owned layer of the rank above: staged from ITS buffers (peer memory when on another GPU)
            MCK(m, cudaStreamWaitEvent(m->stream[r], m->ev_ready[r + 1], 0));
            a.f0_top = vol_two[r + 1];
            a.f1_top = d_result ? d_result[r + 1] : nullptr;
            a.gp_top = vol_topo ? (const GridPoint*)vol_topo[r + 1] : nullptr;
        }
        a.pos = (float4*)pos[r]; a.norm = (float4*)norm[r];
        a.max_verts = max_verts[r];
        a.comp = d_compVoxelArray ? d_compVoxelArray[r] : nullptr;
        a.gz0 = z0[r];   // compacted cell ids are global
        a.gnz = gnz;
        unsigned long long dummy_a = 0, dummy_v = 0;
        if (launch_extract(C, a, &dummy_a, &dummy_v, m->h_totals[r]))
            return mfail(m, std::string("multi_computeIsosurface_2 rank ") + std::to_string(r) + ": " + gcb_last_error(m->ctx[r]));
        MCK(m, cudaEventRecord(m->ev_t1[r], m->stream[r]));
    }
    (void)layer;
    return finish(m, active, verts, vert_offsets, active_offsets);
}

}  // extern "C"
