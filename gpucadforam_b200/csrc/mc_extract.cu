// mc_extract.cu -- fused marching-cubes extraction for sm_100a.
//
// ONE persistent kernel replaces the reference's classify -> thrust scan -> D2H -> compact ->
// thrust scan -> D2H -> memset -> generate sequence (src/Isosurface.cu:44-134 and its four
// siblings; kernels src/MarchingCubes_kernel.cu:869-1063, :1427-1511, :1594-1608, :1865-2200,
// :2625-3017, :3207-3249, :3424-3817).
//
// Work decomposition.  The reference orders output vertices by ascending linear cell id
// (x fastest, then y, then z; `index = numVertsScanned[voxel] + j`, MarchingCubes_kernel.cu:2160).
// A tile is R consecutive y-rows of one z-slice = a CONTIGUOUS range of R*(Nx-1) linear cell ids,
// and tiles are numbered in the same linear order.  Per tile:
//   1. stage-in : the interpolation field rows [y0, y0+R] of point slices z and z+1 are copied
//                 global -> shared with TMA bulk copies (cp.async.bulk + mbarrier); a warp turns 128
//                 staged points at a time into values + inside bits, and the bits leave the warp as
//                 four ballot words = a contiguous bit mask of the row (16 bytes per 128 points);
//   2. classify : a warp classifies 128 cells of a row per step, four consecutive cells per lane:
//                 five mask bits per (slice, row) come from one funnel shift, the four cube indices
//                 are assembled as the bytes of one word (bit spreading by multiply), a step whose
//                 masks are all-0 / all-1 is skipped after the mask test (empty space costs ~25
//                 instructions per 128 cells); vertex counts from a shared-memory table;
//   3. look-back: single-pass decoupled look-back over tiles publishes {active, vertex} prefixes,
//                 so offsets are exactly the reference's exclusive scans;
//   4. emit     : non-empty steps scan their triangle counts (one packed warp scan per 128 cells),
//                 append their triangles to a per-warp ring; the warp drains the ring 32 triangles at
//                 a time (one triangle per lane, vertices interpolated from the STAGED field) and
//                 writes float4 pos/norm at the reference's indices.
// The field is read from HBM once; no per-cell scratch arrays are written (the reference moves
// 36-56 B/cell through them).
#include "common.cuh"

#include <cmath>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

namespace gcb {

// ---------------------------------------------------------------- tables
// Bourke triTable, 16 nibbles per case, 0xF terminator (tools/pack_mc_tables.py).
// Reference: src/tables.h:49-307; numVertsTable (:311-569) is derived = popcount of used nibbles.
__constant__ unsigned long long c_tri_packed[256] = {
#include "mc_tables_packed.inc"
};
static const unsigned long long h_tri_packed[256] = {
#include "mc_tables_packed.inc"
};

void host_tables(unsigned int* tri, unsigned int* nverts) {
    for (int c = 0; c < 256; ++c) {
        unsigned n = 0;
        for (int j = 0; j < 16; ++j) {
            unsigned e = (unsigned)((h_tri_packed[c] >> (4 * j)) & 15ull);
            if (tri) tri[c * 16 + j] = (e == 15u) ? 255u : e;
            if (e != 15u) ++n;
        }
        if (nverts) nverts[c] = n;
    }
}

// ---------------------------------------------------------------- PTX helpers (TMA bulk copy + mbarrier)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---------------------------------------------------------------- constants
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kQueue = 256;  // per-warp triangle ring (power of two, >= 31 + 32*5)
constexpr unsigned long long kFlagAgg = 1ull << 62, kFlagIncl = 2ull << 62, kValMask = (1ull << 62) - 1;

// corner c -> (dx,dy,dz) as 3 bits; MarchingCubes_kernel.cu:889-896
__device__ __forceinline__ uint32_t corner_bits(uint32_t c) {
    // c: 0(000) 1(100) 2(110) 3(010) 4(001) 5(101) 6(111) 7(011)   bits: x=1,y=2,z=4
    return (0x67542310u >> (4 * c)) & 7u;
}
// edge -> endpoints (a | b<<4).  lattice variants: :3751-3762 ; owner variants: :2140-2151
__device__ __forceinline__ uint32_t edge_lat(uint32_t e) {
    // a: 0 1 2 3 4 5 6 7 0 1 2 3   b: 1 2 3 0 5 6 7 4 4 5 6 7
    const unsigned long long A = 0x321076543210ull, B = 0x765447650321ull;
    return (uint32_t)((A >> (4 * e)) & 15ull) | ((uint32_t)((B >> (4 * e)) & 15ull) << 4);
}
__device__ __forceinline__ uint32_t edge_own(uint32_t e) {
    // a: 0 1 3 0 4 5 7 4 0 1 2 3   b: 1 2 2 3 5 6 6 7 4 5 6 7   (first endpoint owns t_x/t_y/t_z)
    const unsigned long long A = 0x321047540310ull, B = 0x765476653221ull;
    return (uint32_t)((A >> (4 * e)) & 15ull) | ((uint32_t)((B >> (4 * e)) & 15ull) << 4);
}
// axis of the stored crossing parameter used by edge e: 0=t_x 1=t_y 2=t_z
__device__ __forceinline__ uint32_t edge_axis(uint32_t e) { return e >= 8 ? 2u : (e & 1u); }

struct Smem {
    float* val[2];
    unsigned char* idb[2];  // lattice modes: id class of the mask value per staged point (0:==0, 1:==1, 2:==2, 3:other)
    uint32_t* plane;        // inside bits: [plane][slice][staged row][128-point chunk][4 words], bit x%128 of a row chunk
    unsigned char* cube;    // cube index per cell, [row][cstride]
    unsigned char* cls;     // M_REGION: which cascade level produced the cube index (0: vol_topo, 1: second test, 2: third test)
    unsigned long long* tri;
    unsigned short* nv;     // numVertsTable[c] | (numVertsTable[c] > 0) << 8
    uint32_t* queue;      // kWarps * kQueue
    uint32_t* warp_tot;   // kWarps * 2
    unsigned long long* prefix;  // [0]=active prefix, [1]=vertex prefix (exclusive, for this tile)
    uint32_t* tile_id;
    uint64_t* mbar;
    uint2* edge;          // [0..11] lattice edge order, [16..27] owner edge order.  x: corner bits ca | cb << 3;
                          // y: staged-point offsets of the two endpoints from the cell's (x, y, z) point, a | b << 16 (slice z+1 = + prow_stride)
};

__device__ __forceinline__ float3 lerp3(float3 a, float3 b, float t) {
    // commons/helper_math.h:1145-1148 `a + t*(b-a)`; the reference build (sm_100a SASS) evaluates it as
    // FADD d=b-a ; FFMA d*t+a.  Spelled with intrinsics so no other contraction can be chosen here.
    return make_float3(__fmaf_rn(__fsub_rn(b.x, a.x), t, a.x), __fmaf_rn(__fsub_rn(b.y, a.y), t, a.y), __fmaf_rn(__fsub_rn(b.z, a.z), t, a.z));
}
__device__ __forceinline__ float3 sub3(float3 a, float3 b) { return make_float3(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)); }
__device__ __forceinline__ float3 cross3(float3 a, float3 b) {  // helper_math.h:1427-1430
    // reference SASS: FMUL second product, FFMA first product minus it
    return make_float3(__fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y)), __fmaf_rn(a.z, b.x, -__fmul_rn(a.x, b.z)), __fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x)));
}

// vertexInterp2_new (MarchingCubes_kernel.cu:3593-3674); t = 0 where the reference leaves it unset
// `fabs(d) < 0.0005` in the reference compares a float against a double literal; kSnap is the smallest float whose value is
// >= 0.0005, so `fabsf(d) < kSnap` decides identically for every float d without leaving the FP32 pipe.
__device__ __forceinline__ bool snap(float d, float thr) { return fabsf(d) < thr; }
// Cell edges are axis aligned: the endpoints p0, p1 the reference hands to vertexInterp2_new differ in ONE coordinate.  Its
// lerp `a + t*(b-a)` leaves the other two untouched (a + t*0 == a for every finite t and a != -0; positions are never -0),
// the snap branches return an endpoint as a whole, and the only non-finite t this function can produce (0/0 when lo == hi ==
// level) is overridden by the first snap test.  So only the coordinate ALONG the edge is interpolated: a / b are that
// coordinate at the two endpoints, za / zb their z coordinates (the reference's `p0.z == 0` special case looks at z).
// Branch-free (lanes of a warp hold unrelated edges): all candidates are computed, the reference's decision tree only selects.
// CHECK_IDS: the id test of the reference (mask ids {1,0}); M_BAND_RAW builds the mask in-kernel from the band test, there an
// edge named by the triangle table joins a point with mask 0 to a point with mask 1 (the cube bits ARE "mask < iso" and the
// mask only takes the values 0 and 1), so the test is true for every edge that gets here and the ids are not even staged.
template <bool CHECK_IDS>
__device__ __forceinline__ float interp_band_axis(float thr, float l1, float l2, float a, float b, float za, float zb, float f0, float f1, uint32_t id0,
                                                  uint32_t id1) {
    const bool crossing = !CHECK_IDS || (id0 == 1u && id1 == 0u) || (id0 == 0u && id1 == 1u);
    const bool sw = f1 < f0;
    const float lo = sw ? f1 : f0, hi = sw ? f0 : f1;
    const float clo = sw ? b : a, chi = sw ? a : b, zlo = sw ? zb : za;
    const bool c1 = (hi >= l1) && (lo <= l1);
    const bool c2 = !c1 && (hi >= l2) && (lo <= l2);
    const float lv = c1 ? l1 : l2;
    const float dn = __fsub_rn(lv, lo), dd = __fsub_rn(hi, lo);
    const bool s_lo = snap(dn, thr) || (!snap(__fsub_rn(lv, hi), thr) && snap(dd, thr));  // -> p0 (1st or 3rd test)
    const bool s_hi = !snap(dn, thr) && snap(__fsub_rn(lv, hi), thr);                        // -> p1 (2nd test)
    float t = __fdiv_rn(dn, dd);
    if (!(c1 || c2)) t = (hi == lo && zlo == 0.0f) ? 1.f : 0.f;  // reference: t = 1 / t = 0 / (unset -> 0 here)
    if (!crossing) t = 0.f;
    // without a crossing the reference lerps the UNSWAPPED endpoints with t = 0, i.e. returns p0 + 0*(p1-p0)
    const float ea = crossing ? clo : a, eb = crossing ? chi : b;
    float r = __fmaf_rn(__fsub_rn(eb, ea), t, ea);  // lerp, helper_math.h:1145-1148 as the reference build contracts it
    if (crossing && (c1 || c2)) {
        if (s_lo) r = clo;
        else if (s_hi) r = chi;
    }
    return r;
}
// second half of vertexInterp3_new (:3347-3413): band on the vol_two pair when ids are {2,0}
__device__ __forceinline__ bool interp_band_two(float thr, float m1, float m2, float3& p0, float3& p1, float f2, float f3, float& t, float3& out) {
    if (f3 < f2) { float3 tp = p1; p1 = p0; p0 = tp; float tf = f3; f3 = f2; f2 = tf; }
    if ((f3 >= m1) && (f2 <= m1)) {
        if (snap(__fsub_rn(m1, f2), thr)) { out = p0; return true; }
        if (snap(__fsub_rn(m1, f3), thr)) { out = p1; return true; }
        if (snap(__fsub_rn(f3, f2), thr)) { out = p0; return true; }
        t = __fdiv_rn(__fsub_rn(m1, f2), __fsub_rn(f3, f2));
    } else if ((f3 >= m2) && (f2 <= m2)) {
        if (snap(__fsub_rn(m2, f2), thr)) { out = p0; return true; }
        if (snap(__fsub_rn(m2, f3), thr)) { out = p1; return true; }
        if (snap(__fsub_rn(f3, f2), thr)) { out = p0; return true; }
        t = __fdiv_rn(__fsub_rn(m2, f2), __fsub_rn(f3, f2));
    } else if ((f3 == f2) && (p0.z == 0.0f)) t = 1;
    else if (f3 == f2) t = 0;
    return false;
}
__device__ __forceinline__ float blend_t(float t1, float t2, float t) {  // :1657-1668
    if ((t1 > 0) && (t2 > 0)) t = (t1 + t2) * 0.5;
    else if ((t1 > 0) && (t2 == 0)) t = t1;
    else if ((t2 > 0) && (t1 == 0)) t = t2;
    return t;
}
__device__ __forceinline__ float t_primitive(float iso, float f0, float f1, float et) {  // :1640-1672
    float t2 = 0.0;
    if (((f1 >= iso) && (f0 <= iso)) || ((f0 >= iso) && (f1 <= iso))) t2 = __fdiv_rn(__fsub_rn(iso, f0), __fsub_rn(f1, f0));
    return blend_t(et, t2, 0);
}
__device__ __forceinline__ float t_primitive_one(float l1, float l2, float f0, float f1, float et) {  // :1675-1724
    float t2 = 0.0, t3 = 0.0;
    if (((f1 >= l1) && (f0 <= l1)) || ((f0 >= l1) && (f1 <= l1))) t2 = __fdiv_rn(__fsub_rn(l1, f0), __fsub_rn(f1, f0));
    float t = blend_t(et, t2, 0);
    if (((f1 >= l2) && (f0 <= l2)) || ((f0 >= l2) && (f1 <= l2))) t3 = __fdiv_rn(__fsub_rn(l2, f0), __fsub_rn(f1, f0));
    return blend_t(et, t3, t);
}
__device__ __forceinline__ float t_fixed(float iso, float l1, float l2, float f0, float f1, float f2, float f3) {  // :1727-1784
    float t1 = 0.0f, t2 = 0.0f, t3 = 0.0f, t = 0.0f;
    if (((f0 < iso) && (f1 >= iso)) || ((f1 < iso) && (f0 >= iso))) t1 = __fdiv_rn(__fsub_rn(iso, f0), __fsub_rn(f1, f0));
    if (((f2 < l1) && (f3 >= l1)) || ((f3 < l1) && (f2 >= l1))) t2 = __fdiv_rn(__fsub_rn(l1, f2), __fsub_rn(f3, f2));
    if (((f2 < l2) && (f3 >= l2)) || ((f3 < l2) && (f2 >= l2))) t3 = __fdiv_rn(__fsub_rn(l2, f2), __fsub_rn(f3, f2));
    if ((t1 > 0.0f) && (t2 > 0.0f) && (t3 == 0.0)) t = (t1 + t2) * 0.5;
    else if ((t1 > 0.0f) && (t3 > 0.0f) && (t2 == 0.0f)) t = (t1 + t3) * 0.5;
    else if ((t1 > 0.0) && (t2 == 0.0) && (t3 == 0.0)) t = t1;
    else if ((t2 > 0.0) && (t1 == 0.0) && (t3 == 0.0)) t = t2;
    else if ((t3 > 0.0) && (t1 == 0.0) && (t2 == 0.0)) t = t3;
    return t;
}
__device__ __forceinline__ float t_analysis(float iso, float f0, float f1, float et) {  // :1787-1820
    float t2 = 0.0;
    if (((f1 >= iso) && (f0 < iso)) || ((f0 >= iso) && (f1 < iso))) t2 = __fdiv_rn(__fsub_rn(iso, f0), __fsub_rn(f1, f0));
    return blend_t(et, t2, 0);
}

// ---------------------------------------------------------------- stage-in: one grid point -> {value, bits}
// bits: [1:0] id class of the mask value (0:==0, 1:==1, 2:==2, 3:other), [2] inside flag.
// n / d for a divisor that is the same for every point: y = RN(1/d) is computed once, the quotient is q = n*y followed by two
// remainder corrections (r = n - q*d exactly by fma, q += r*y).  With |d| in [2^-40, 2^40] this equals __fdiv_rn(n, d) bit for
// bit for every n whose remainder cannot underflow; tiny/zero numerators, tiny/huge quotients and inf/nan take the IEEE
// division.  Checked exhaustively (all 2^32 numerators for 118 divisors incl. all-ones mantissas): tools/div_check.cu,
// profiles/r01_div_check.txt.
struct UniformDiv {
    float d, y; bool ok;
    float a;                   // the offset subtracted before the division (M_BAND_RAW: the field minimum)
    bool two; float a2, d2;    // optional second normalisation k <- (k - a2) / d2 (the legacy normalise_buffer + normalise_four sequence)
};
__device__ __forceinline__ UniformDiv make_uniform_div(float d) {
    UniformDiv u;
    u.d = d;
    u.y = __frcp_rn(d);
    u.ok = fabsf(d) >= 0x1p-40f && fabsf(d) <= 0x1p40f;
    return u;
}
__device__ __forceinline__ float div_by_uniform(float n, const UniformDiv& u) {
    const float q0 = __fmul_rn(n, u.y);
    const float r0 = __fmaf_rn(-q0, u.d, n);
    const float q1 = __fmaf_rn(r0, u.y, q0);
    const float r1 = __fmaf_rn(-q1, u.d, n);
    const float q2 = __fmaf_rn(r1, u.y, q1);
    if (!(u.ok && fabsf(n) >= 1.0e-30f && fabsf(q2) >= 1.0e-30f && fabsf(q2) <= 1.0e30f)) return __fdiv_rn(n, u.d);
    return q2;
}

// Stored fields under z-slab sharding: a rank may hold only its OWNED point layers, the +z halo layer (local layer nz - 1) then
// lives in the upper neighbour's memory (`*_top` = start of that layer, a peer-mapped pointer over NVLink or an ordinary one).
template <class T>
__device__ __forceinline__ const T* at_point(const T* base, const T* top, size_t gi, size_t top_begin) {
    return (top && gi >= top_begin) ? top + (gi - top_begin) : base + gi;
}
// What one point needs from global memory besides the TMA-staged field: fetched ahead of use (stage_fetch), turned into
// {value, bits} later (stage_point), so a warp keeps the loads of its next 128 points in flight while it evaluates the current ones.
template <int MODE>
__device__ __forceinline__ void stage_fetch(const McArgs& A, size_t gi, float& a, float& b) {
    if (MODE == M_LATTICE_ONE || MODE == M_LATTICE) a = __ldg(at_point(A.f1, A.f1_top, gi, A.top_begin));  // mask `vol`
    else if (MODE == M_REGION) {
        a = A.gp2 ? (float)A.gp2[gi].val : 0.f;  // sampleVolume_2 :98-108
        b = A.gp ? (float)A.gp[gi].val : 0.f;
    } else if (MODE == M_TOPO) a = A.gp ? (float)at_point(A.gp, A.gp_top, gi, A.top_begin)->val : 0.f;
    else if (MODE == M_CSG) {
        a = A.gp ? (float)at_point(A.gp, A.gp_top, gi, A.top_begin)->val : 0.f;
        if ((A.flags & (F_FIXED | F_DYNAMIC)) && A.f1) b = __ldg(at_point(A.f1, A.f1_top, gi, A.top_begin));
    }
}
template <int MODE>
__device__ __forceinline__ void stage_point(const McArgs& A, const UniformDiv& nd, uint32_t x, bool row_face, float raw, float in_a, float in_b, float& val, uint32_t& bits) {
    if (MODE == M_LATTICE_ONE || MODE == M_LATTICE) {
        const float m = in_a;  // classifyVoxel_new :3232-3239
        uint32_t id = (m == 1.f) ? 1u : (m == 0.f) ? 0u : (m == 2.f) ? 2u : 3u;
        bits = id | ((m < A.iso) ? 4u : 0u);
        val = raw;  // k `vol_one`
    } else if (MODE == M_BAND_RAW) {
        // device_bufferfour (Gratings.cu:1089-1134) fused; domain faces use GLOBAL coordinates
        float k = div_by_uniform(__fsub_rn(raw, nd.a), nd);
        if (nd.two) k = __fdiv_rn(__fsub_rn(k, nd.a2), nd.d2);
        float m;
        if (row_face || x == 0 || x == A.nx - 1) { m = 0.0f; k = 0.0f; }
        else m = ((k >= A.iso1) && (k <= A.iso2)) ? 1.0f : 0.0f;
        bits = (m == 1.f ? 1u : 0u) | ((m < A.iso) ? 4u : 0u);
        val = k;
    } else if (MODE == M_REGION) {
        // classifyVoxel_region_kernel :1163-1290: three candidate inside tests per point, the cascade is resolved per cell
        const float iso = A.iso;
        const float ft = in_a, fx = in_b;
        bits = ((ft < iso) ? 4u : 0u) | ((fx < iso) ? 8u : 0u) | (((fx < iso) & (raw < iso)) ? 16u : 0u);
        val = raw;  // primitive_dynamic
    } else if (MODE == M_TOPO) {
        // classifyVoxel_kernel_topo :1492-1499
        const float fx = in_a;
        bits = ((fx < A.iso1) | (raw >= A.iso)) ? 4u : 0u;
        val = raw;
    } else {  // M_CSG  classifyVoxel :922-1052
        const float iso = A.iso;
        const float fx = in_a;
        const float dy = raw;
        const bool fixed = A.flags & F_FIXED, dyn = A.flags & F_DYNAMIC;
        const float la = in_b;
        const bool inb = (la > A.iso1) & (la < A.iso2);
        bool b = false;
        if (A.flags & F_MAKE_REGION) b = fx < iso;
        else if (A.flags & F_UNION) b = fixed ? ((dy < iso) | inb) : dyn ? ((fx < iso) | inb) : ((fx < iso) | (dy < iso));
        else if (A.flags & F_DIFF) b = fixed ? ((dy >= iso) & inb) : dyn ? ((fx < iso) & ((la < A.iso1) | (la > A.iso2))) : ((dy >= iso) & (fx < iso));
        else if (A.flags & F_INTERSECT) b = fixed ? ((dy < iso) & inb) : dyn ? ((fx < iso) & inb) : ((fx < iso) & (dy < iso));
        bits = b ? 4u : 0u;
        val = raw;
    }
}

// ---------------------------------------------------------------- one triangle
// ring entry: triangle number within the cell << 29 | tile row << 16 | x
template <int MODE>
__device__ __forceinline__ void emit_triangle(const McArgs& A, const Smem& S, uint32_t z, uint32_t y0, uint32_t ent, unsigned long long vidx) {
    float3 v[3], n;
    float w[3];
    const uint32_t j = ent >> 29, r = (ent >> 16) & 0x1fffu, x = ent & 0xffffu;
    const uint32_t c = r * A.cx + x;  // cell id within the tile
    const uint32_t cube = S.cube[r * A.cstride + x];
    const unsigned long long tri = S.tri[cube];
    const uint32_t y = y0 + r;
    // MarchingCubes_kernel.cu:1888-1890 : (uint -> float) - center, times voxel
    const float3 p = make_float3(__fmul_rn(__fsub_rn((float)x, A.center.x), A.voxel.x), __fmul_rn(__fsub_rn((float)y, A.center.y), A.voxel.y),
                                 __fmul_rn(__fsub_rn((float)(z + A.gz0), A.center.z), A.voxel.z));  // global z under slab sharding

    const float3 pmax = make_float3(__fadd_rn(p.x, A.voxel.x), __fadd_rn(p.y, A.voxel.y), __fadd_rn(p.z, A.voxel.z));
    const uint32_t sbase = r * A.nx + x;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const uint32_t e = (uint32_t)(tri >> (4 * (3 * j + k))) & 15u;
        const bool own = (MODE == M_TOPO) || (MODE == M_REGION) || (MODE == M_CSG && !(A.flags & F_FIXED));
        const uint2 ed = S.edge[(own ? 16u : 0u) + e];
        const uint32_t ca = ed.x & 7u, cb = ed.x >> 3;
        // corner positions: v[0] = p, v[i] = p + (voxel or 0) per component (:1892-1900).  p + 0.0f == p bit for bit for every
        // p this kernel can produce (x - center is never -0, voxel sizes are positive), so each component is a select.
        const float3 pa = make_float3((ca & 1u) ? pmax.x : p.x, (ca & 2u) ? pmax.y : p.y, (ca & 4u) ? pmax.z : p.z);
        const uint32_t sa = sbase + (ed.y & 0xffffu), sb = sbase + (ed.y >> 16);  // both staged slices are one array: val[1] == val[0] + prow_stride
        const uint32_t za = (ca >> 2) & 1u, zb = (cb >> 2) & 1u;
        const float fa = S.val[0][sa], fb = S.val[0][sb];
        w[k] = 0.f;
        if (MODE == M_BAND_RAW || MODE == M_LATTICE_ONE || MODE == M_LATTICE) {
            // band interpolation along the edge's axis (see interp_band_axis)
            const uint32_t ax = ca ^ cb;  // 1: x, 2: y, 4: z
            const float c0 = ax == 1u ? p.x : ax == 2u ? p.y : p.z, c1 = ax == 1u ? pmax.x : ax == 2u ? pmax.y : pmax.z;
            const float ea = (ca & ax) ? c1 : c0, eb = (cb & ax) ? c1 : c0;
            const float pzb = (cb & 4u) ? pmax.z : p.z;
            uint32_t ida = 0, idb = 0;
            if (MODE != M_BAND_RAW) { ida = S.idb[0][sa]; idb = S.idb[0][sb]; }  // idb[1] == idb[0] + prow_stride
            bool done = false;
            if (MODE == M_LATTICE) {
                if ((ida == 2u && idb == 0u) || (idb == 2u && ida == 0u)) {
                    // ids {2,0}: first block of vertexInterp3_new does not fire, second one does
                    const size_t ga = ((size_t)(z + za) * A.ny + y + ((ca >> 1) & 1u)) * A.nx + x + (ca & 1u);
                    const size_t gb = ((size_t)(z + zb) * A.ny + y + ((cb >> 1) & 1u)) * A.nx + x + (cb & 1u);
                    const float3 pb = make_float3((cb & 1u) ? pmax.x : p.x, (cb & 2u) ? pmax.y : p.y, (cb & 4u) ? pmax.z : p.z);
                    float t = 0.f;
                    float3 out;
                    float3 q0 = pa, q1 = pb;
                    if (interp_band_two(A.snap_thr, A.iso1b, A.iso2b, q0, q1, __ldg(A.f2 + ga), __ldg(A.f2 + gb), t, out)) v[k] = out;
                    else v[k] = lerp3(q0, q1, t);
                    done = true;
                }
            }
            if (!done) {
                const float rr = (MODE == M_BAND_RAW) ? interp_band_axis<false>(A.snap_thr, A.iso1, A.iso2, ea, eb, pa.z, pzb, fa, fb, 0u, 0u)
                                                      : interp_band_axis<true>(A.snap_thr, A.iso1, A.iso2, ea, eb, pa.z, pzb, fa, fb, ida, idb);
                v[k] = make_float3(ax == 1u ? rr : pa.x, ax == 2u ? rr : pa.y, ax == 4u ? rr : pa.z);
            }
            continue;
        }
        const float3 pb = make_float3((cb & 1u) ? pmax.x : p.x, (cb & 2u) ? pmax.y : p.y, (cb & 4u) ? pmax.z : p.z);
        if (MODE == M_REGION) {
            // generateTriangles_region_kernel :2391-2490: stored crossing parameter of the edge's owning point, from vol_topo
            // (first test fired) or primitive_fixed; make_region's second test blends it with the dynamic field's crossing
            const size_t ga = ((size_t)(z + za) * A.ny + y + ((ca >> 1) & 1u)) * A.nx + x + (ca & 1u);
            const uint32_t cl = S.cls[r * A.cstride + x];
            const GridPoint* src = cl == 0u ? A.gp2 : A.gp;
            float et = 0.f;
            if (src) {
                const GridPoint g = src[ga];
                const uint32_t ax = edge_axis(e);
                et = ax == 0 ? g.t_x : ax == 1 ? g.t_y : g.t_z;
            }
            const float t = (cl == 1u && !(A.flags & F_SHOW_DOMAIN)) ? t_primitive(A.iso, fa, fb, et) : et;
            v[k] = lerp3(pa, pb, t);
            w[k] = cl == 0u ? 1.0f : cl == 1u ? 0.25f : 0.5f;  // `aa`
        } else {
            const size_t ga = ((size_t)(z + za) * A.ny + y + ((ca >> 1) & 1u)) * A.nx + x + (ca & 1u);
            const size_t gb = ((size_t)(z + zb) * A.ny + y + ((cb >> 1) & 1u)) * A.nx + x + (cb & 1u);
            float et = 0.f;
            if (own && A.gp) {
                const GridPoint g = *at_point(A.gp, A.gp_top, ga, A.top_begin);
                const uint32_t ax = edge_axis(e);
                et = ax == 0 ? g.t_x : ax == 1 ? g.t_y : g.t_z;
            }
            float t;
            float3 qa = pa, qb = pb;
            if (MODE == M_TOPO) {
                t = t_analysis(A.iso, fa, fb, et);
                w[k] = A.f1 ? __ldg(at_point(A.f1, A.f1_top, ga, A.top_begin)) : 0.f;  // *field_val = r0 (:1797)
                if (A.flags & F_DISP) {
                    const float4 d0 = A.disp[ga], d1 = A.disp[gb];
                    qa = make_float3(d0.x, d0.y, d0.z);
                    qb = make_float3(d1.x, d1.y, d1.z);
                }
            } else if (A.flags & F_MAKE_REGION) t = et;
            else if (A.flags & F_FIXED)
                t = t_fixed(A.iso, A.iso1, A.iso2, fa, fb, __ldg(at_point(A.f1, A.f1_top, ga, A.top_begin)), __ldg(at_point(A.f1, A.f1_top, gb, A.top_begin)));
            else if (A.flags & F_DYNAMIC)
                t = t_primitive_one(A.iso1, A.iso2, __ldg(at_point(A.f1, A.f1_top, ga, A.top_begin)), __ldg(at_point(A.f1, A.f1_top, gb, A.top_begin)), et);
            else t = t_primitive(A.iso, fa, fb, et);
            v[k] = lerp3(qa, qb, t);
        }
    }
    if (MODE == M_CSG) {  // calcNormal(ver0, ver2, ver1), w = 0.5 (:2178-2185)
        n = cross3(sub3(v[2], v[0]), sub3(v[1], v[0]));
        w[0] = w[1] = w[2] = 0.5f;
    } else if (MODE == M_REGION) {
        // normalize(calcNormal(v0, v1, v2)) :2576; helper_math.h normalize = v * rsqrtf(dot(v, v)).  The reference build
        // (sm_100a SASS) evaluates the dot product as FMUL y*y, FFMA x*x + ., FFMA z*z + . and rsqrtf through MUFU.RSQ with
        // the denormal pre/post scaling -- rsqrtf() here compiles to the same sequence.
        n = cross3(sub3(v[1], v[0]), sub3(v[2], v[0]));
        const float inv = rsqrtf(__fmaf_rn(n.z, n.z, __fmaf_rn(n.x, n.x, __fmul_rn(n.y, n.y))));
        n = make_float3(__fmul_rn(n.x, inv), __fmul_rn(n.y, inv), __fmul_rn(n.z, inv));
        if ((A.flags & F_SHOW_REGION) && A.meta) {
            // triangle_metadata :2528-2584, written whether or not the vertices fit into maxVerts; load_group / force_dir untouched
            const uint32_t ind = (uint32_t)(vidx / 3ull);
            TriangleMetadata* m = A.meta + ind;
            m->index = ind;
            m->voxel = (z + A.gz0) * A.cx * A.cy + y0 * A.cx + c;
            m->l_index = j;
            m->edge_1 = (uint32_t)(tri >> (12 * j)) & 15u;
            m->edge_2 = (uint32_t)(tri >> (12 * j + 4)) & 15u;
            m->edge_3 = (uint32_t)(tri >> (12 * j + 8)) & 15u;
            m->centroid[0] = __fdiv_rn(__fadd_rn(__fadd_rn(v[0].x, v[1].x), v[2].x), 3.0f);
            m->centroid[1] = __fdiv_rn(__fadd_rn(__fadd_rn(v[0].y, v[1].y), v[2].y), 3.0f);
            m->centroid[2] = __fdiv_rn(__fadd_rn(__fadd_rn(v[0].z, v[1].z), v[2].z), 3.0f);
            m->normal[0] = n.x; m->normal[1] = n.y; m->normal[2] = n.z;
        }
    } else {
        n = cross3(sub3(v[1], v[0]), sub3(v[2], v[0]));
    }
    // `index < maxVerts - 3` in unsigned arithmetic (:2181).  For maxVerts < 3 that expression wraps to ~4e9 in the reference and
    // every triangle would be written past the end of the caller's buffers; this library writes nothing then (counts are still
    // reported).  Capacities above 2^32 - 1 (fused entry points only) use the exact test `index + 3 <= maxVerts`.
    const unsigned long long limit = A.max_verts < 3ull ? 0ull : (unsigned long long)((unsigned int)A.max_verts - 3u);
    const bool ok = (A.max_verts > 0xffffffffull) ? (vidx + 3 <= A.max_verts) : (vidx < limit);
    if (ok) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            __stcs(A.pos + vidx + k, make_float4(v[k].x, v[k].y, v[k].z, 1.0f));
            __stcs(A.norm + vidx + k, make_float4(n.x, n.y, n.z, w[k]));
        }
    }
}

// ---------------------------------------------------------------- the kernel
template <int MODE> struct ModeTraits {
    static constexpr int kPlanes = (MODE == M_REGION) ? 3 : 1;                          // inside tests per point
    static constexpr bool kIds = (MODE == M_LATTICE_ONE || MODE == M_LATTICE);          // emission looks at mask ids
#ifndef GCB_MC_MINB
#define GCB_MC_MINB 4
#endif
    static constexpr int kMinBlocks = (MODE == M_BAND_RAW) ? GCB_MC_MINB : 3;  // CTAs per SM the register budget allows
};
// bit u of a 4-bit value -> bit 8u (the four partial products land on distinct bits, so the multiply cannot carry)
__device__ __forceinline__ uint32_t spread4(uint32_t n) { return (n * 0x00204081u) & 0x01010101u; }
// the four cube indices of four consecutive cells of a row as the bytes of one word.  n00/n01: five consecutive inside bits
// (points x..x+4) of rows y / y+1 in slice z, n10/n11 the same in slice z+1.  Corner order MarchingCubes_kernel.cu:889-896.
__device__ __forceinline__ uint32_t cubes4(uint32_t n00, uint32_t n01, uint32_t n10, uint32_t n11) {
    return spread4(n00 & 15u) | (spread4(n00 >> 1) << 1) | (spread4(n01 >> 1) << 2) | (spread4(n01 & 15u) << 3) | (spread4(n10 & 15u) << 4) |
           (spread4(n10 >> 1) << 5) | (spread4(n11 >> 1) << 6) | (spread4(n11 & 15u) << 7);
}

// TMA: the interpolation field is staged by bulk copies (rows 16-byte aligned: nx % 4 == 0); otherwise through LDG.
template <int MODE, bool TMA>
__global__ void __launch_bounds__(kThreads, ModeTraits<MODE>::kMinBlocks) mc_fused_kernel(const McArgs A) {
    constexpr int NP = ModeTraits<MODE>::kPlanes;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem S;
    const uint32_t R1 = A.rows_per_tile + 1;  // staged rows per slice
    {
        unsigned char* p = smem_raw;
        S.val[0] = (float*)p; p += (size_t)A.prow_stride * 4;
        S.val[1] = (float*)p; p += (size_t)A.prow_stride * 4;
        S.tri = (unsigned long long*)p; p += 256 * 8;
        S.prefix = (unsigned long long*)p; p += 16;
        S.mbar = (uint64_t*)p; p += 8;
        S.tile_id = (uint32_t*)p; p += 8;
        S.edge = (uint2*)p; p += 32 * 8;
        S.queue = (uint32_t*)p; p += kWarps * kQueue * 4;
        S.warp_tot = (uint32_t*)p; p += kWarps * 2 * 4;
        S.nv = (unsigned short*)p; p += 512;
        S.plane = (uint32_t*)p; p += (size_t)NP * 2 * R1 * A.ppr * 16 + 16;
        S.cube = p; p += (size_t)A.rows_per_tile * A.cstride;
        S.cls = p; if (MODE == M_REGION) p += (size_t)A.rows_per_tile * A.cstride;
        S.idb[0] = p; if (ModeTraits<MODE>::kIds) p += A.prow_stride;
        S.idb[1] = p;
    }
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

    for (uint32_t i = tid; i < 256; i += kThreads) {
        const unsigned long long t = c_tri_packed[i];
        S.tri[i] = t;
        // number of used nibbles = numVertsTable[i]
        unsigned long long u = ~t;                      // used nibble != 0xF  <=> ~nibble != 0
        u = (u | (u >> 1) | (u >> 2) | (u >> 3)) & 0x1111111111111111ull;
        const uint32_t n = (uint32_t)__popcll(u);
        S.nv[i] = (unsigned short)(n | (n ? 256u : 0u));
    }
    if (tid < 12) {
        const uint32_t l = edge_lat(tid), o = edge_own(tid);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const uint32_t ab = t ? o : l, ca = corner_bits(ab & 15u), cb = corner_bits(ab >> 4);
            const uint32_t oa = ((ca & 4u) ? A.prow_stride : 0u) + ((ca & 2u) ? A.nx : 0u) + (ca & 1u);
            const uint32_t ob = ((cb & 4u) ? A.prow_stride : 0u) + ((cb & 2u) ? A.nx : 0u) + (cb & 1u);
            S.edge[16 * t + tid] = make_uint2(ca | (cb << 3), oa | (ob << 16));
        }
    }
    if (tid == 0) { mbar_init(S.mbar, 1); fence_mbar_init(); }
    __syncthreads();

    uint32_t parity = 0;
    const uint32_t slice_pts = A.nx * A.ny;
    // M_BAND_RAW normalisation range {min, max}: kernel arguments, or read from device memory where the field kernel's reduction
    // left it (no host round trip between field and extraction)
    float na = A.na, nb = A.nb;
    if (MODE == M_BAND_RAW && A.d_ab) { na = __ldg(A.d_ab); nb = __ldg(A.d_ab + 1); }
    UniformDiv nd = make_uniform_div(__fsub_rn(nb, na));
    nd.a = na;
    nd.two = false; nd.a2 = 0.f; nd.d2 = 1.f;
    if (MODE == M_BAND_RAW && A.d_ab && A.two_stage) {
        // {a2, b2}: range of the once-normalised field (d_ab[2..3]); (k - 0) / 1 is the identity, the usual case
        const float a2 = __ldg(A.d_ab + 2), b2 = __ldg(A.d_ab + 3);
        nd.two = !(a2 == 0.f && b2 == 1.f);
        nd.a2 = a2; nd.d2 = __fsub_rn(b2, a2);
    }
    const uint32_t ppr = A.ppr, cpr = A.cpr;
    const uint32_t plane_words = 2u * R1 * ppr * 4u;  // words per bit plane
    constexpr bool tma = TMA;

    for (;;) {
        if (tid == 0) *S.tile_id = atomicAdd(A.tile_counter, 1u);
        __syncthreads();  // also orders the previous tile's smem reads before this tile's writes
        const uint32_t tile = *S.tile_id;
        if (tile >= A.num_tiles) break;
        const uint32_t z = tile / A.tiles_per_slice, ty = tile - z * A.tiles_per_slice;
        const uint32_t y0 = ty * A.rows_per_tile;
        const uint32_t rows = min(A.rows_per_tile, A.cy - y0);
        const uint32_t npts = (rows + 1) * A.nx;   // staged points per slice
        const size_t g0 = (size_t)z * slice_pts + (size_t)y0 * A.nx;  // first staged point, slice z
        const bool top_peer = A.f0_top != nullptr && z + 2u == A.nz;    // slice z + 1 is the +z halo layer held by the upper neighbour

        // ---- 1. stage-in
        if (TMA) {
            if (tid == 0) {
                fence_proxy_async();  // generic-proxy accesses of the previous tile precede the async writes
                mbar_expect_tx(S.mbar, 2u * npts * 4u);
                tma_bulk_g2s(S.val[0], A.f0 + g0, npts * 4u, S.mbar);
                // slice z + 1: the rank's own layer, or -- for the last cell layer of a slab that holds owned layers only -- the first
                // owned layer of the upper neighbour, read where it lies (peer memory over NVLink): no halo copy, no exchange step
                tma_bulk_g2s(S.val[1], top_peer ? A.f0_top + (size_t)y0 * A.nx : A.f0 + g0 + slice_pts, npts * 4u, S.mbar);
            }
            mbar_wait(S.mbar, parity);
            parity ^= 1u;
        }
        if (TMA && MODE != M_REGION) {
            // Field staged by TMA (rows 16-byte aligned in shared memory, nx % 4 == 0).  Work item = 128 points of one staged row,
            // dealt round-robin to the warps; lane l takes points 4l .. 4l+3 (one LDS.128), its four inside bits form a nibble, and
            // an OR over each group of eight lanes assembles the 32-point words of the row's bit mask.  M_BAND_RAW (the hot
            // configuration) fuses normalisation + domain faces + band test here and writes k back with one STS.128; the other
            // modes fetch what else a point needs from global memory (four independent loads per lane) and leave the values alone.
            const uint32_t lim = 2u * (rows + 1u);
            const bool in0 = 0.f < A.iso, in1 = 1.f < A.iso;  // M_BAND_RAW: "mask < iso" for the two values the mask takes
            uint32_t k = warp % ppr, rs = warp / ppr;         // rs: staged row over both slices, [0, lim)
            while (rs < lim) {
                const uint32_t s = rs > rows ? 1u : 0u, rr = rs - s * (rows + 1u);
                float* sv = (s ? S.val[1] : S.val[0]) + rr * A.nx;
                const uint32_t yy = y0 + rr, gz = z + s + A.gz0;
                const bool row_face = yy == 0 || yy == A.ny - 1 || gz == 0 || gz == A.gnz - 1;  // domain faces in GLOBAL coordinates
                const uint32_t x = 128u * k + 4u * lane;
                uint32_t nib = 0;
                if (x < A.nx) {
                    const float4 v4 = *reinterpret_cast<const float4*>(sv + x);
                    if (MODE == M_BAND_RAW) {
                        float nn[4] = {__fsub_rn(v4.x, na), __fsub_rn(v4.y, na), __fsub_rn(v4.z, na), __fsub_rn(v4.w, na)}, kk[4];
                        bool fast = nd.ok;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {  // div_by_uniform, the four range checks folded into one branch
                            const float q0 = __fmul_rn(nn[u], nd.y);
                            const float r0 = __fmaf_rn(-q0, nd.d, nn[u]);
                            const float q1 = __fmaf_rn(r0, nd.y, q0);
                            const float r1 = __fmaf_rn(-q1, nd.d, nn[u]);
                            kk[u] = __fmaf_rn(r1, nd.y, q1);
                            fast = fast && fabsf(nn[u]) >= 1.0e-30f && fabsf(kk[u]) >= 1.0e-30f && fabsf(kk[u]) <= 1.0e30f;
                        }
                        if (!fast) {
#pragma unroll
                            for (int u = 0; u < 4; ++u) kk[u] = div_by_uniform(nn[u], nd);
                        }
                        if (nd.two) {
#pragma unroll
                            for (int u = 0; u < 4; ++u) kk[u] = __fdiv_rn(__fsub_rn(kk[u], nd.a2), nd.d2);
                        }
                        // device_bufferfour (Gratings.cu:1089-1134): band mask, the six domain faces forced to k = 0, m = 0
                        bool m[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) m[u] = (kk[u] >= A.iso1) && (kk[u] <= A.iso2);
                        if (row_face) {
#pragma unroll
                            for (int u = 0; u < 4; ++u) { kk[u] = 0.f; m[u] = false; }
                        }
                        if (x == 0u) { kk[0] = 0.f; m[0] = false; }
                        if (x + 4u == A.nx) { kk[3] = 0.f; m[3] = false; }
#pragma unroll
                        for (int u = 0; u < 4; ++u) nib |= ((m[u] ? in1 : in0) ? 1u : 0u) << u;
                        *reinterpret_cast<float4*>(sv + x) = make_float4(kk[0], kk[1], kk[2], kk[3]);
                    } else {
                        const size_t gi = g0 + (size_t)s * slice_pts + (size_t)rr * A.nx + x;
                        const float raw[4] = {v4.x, v4.y, v4.z, v4.w};
                        float fa[4] = {0.f, 0.f, 0.f, 0.f}, fb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int u = 0; u < 4; ++u) stage_fetch<MODE>(A, gi + u, fa[u], fb[u]);
                        uint32_t idw = 0;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            float val;
                            uint32_t bits;
                            stage_point<MODE>(A, nd, x + u, row_face, raw[u], fa[u], fb[u], val, bits);
                            nib |= ((bits >> 2) & 1u) << u;
                            idw |= (bits & 3u) << (8 * u);
                        }
                        if (ModeTraits<MODE>::kIds) *reinterpret_cast<uint32_t*>((s ? S.idb[1] : S.idb[0]) + rr * A.nx + x) = idw;
                    }
                }
                uint32_t wd = nib << (4u * (lane & 7u));
                wd |= __shfl_xor_sync(0xffffffffu, wd, 1);
                wd |= __shfl_xor_sync(0xffffffffu, wd, 2);
                wd |= __shfl_xor_sync(0xffffffffu, wd, 4);
                if ((lane & 7u) == 0u) S.plane[((s * R1 + rr) * ppr + k) * 4u + (lane >> 3)] = wd;
                k += kWarps;
                while (k >= ppr) { k -= ppr; ++rs; }
            }
        } else {
            // work item = 128 points of one staged row; items are dealt round-robin to the warps.  Lane l takes points
            // 32u + l (u = 0..3) of the chunk, so the four ballots ARE the row's bit mask, 32 points per word.
#ifndef GCB_MC_AHEAD
#define GCB_MC_AHEAD 1
#endif
            constexpr bool kAhead = GCB_MC_AHEAD && MODE != M_BAND_RAW;  // modes with global loads per point fetch one item ahead
            const uint32_t lim = 2u * (rows + 1u);
            uint32_t k = warp % ppr, rs = warp / ppr;  // rs: staged row over both slices, [0, lim)
            float craw[4], ca[4], cb[4], nraw[4], na[4], nb[4];
#define GCB_FETCH(fk, frs, RAW, FA, FB)                                                                     \
    {                                                                                                       \
        const uint32_t s_ = (frs) > rows ? 1u : 0u, rr_ = (frs) - s_ * (rows + 1u);                         \
        const float* sv_ = (s_ ? S.val[1] : S.val[0]) + rr_ * A.nx;                                         \
        const size_t grow_ = g0 + (size_t)s_ * slice_pts + (size_t)rr_ * A.nx;                              \
        _Pragma("unroll") for (int u = 0; u < 4; ++u) {                                                     \
            const uint32_t x_ = 128u * (fk) + 32u * u + lane;                                               \
            RAW[u] = 0.f; FA[u] = 0.f; FB[u] = 0.f;                                                         \
            if (x_ < A.nx) {                                                                                \
                RAW[u] = tma ? sv_[x_] : (A.f0 ? __ldg(((s_ && top_peer) ? A.f0_top + (size_t)(y0 + rr_) * A.nx : A.f0 + grow_) + x_) : 0.f); \
                stage_fetch<MODE>(A, grow_ + x_, FA[u], FB[u]);                                             \
            }                                                                                               \
        }                                                                                                   \
    }
            if (kAhead && rs < lim) GCB_FETCH(k, rs, craw, ca, cb)
            while (rs < lim) {
                uint32_t nk = 0, nrs = 0;
                if (kAhead) {
                    nk = k + kWarps; nrs = rs;
                    while (nk >= ppr) { nk -= ppr; ++nrs; }
                    if (nrs < lim) GCB_FETCH(nk, nrs, nraw, na, nb)
                } else GCB_FETCH(k, rs, craw, ca, cb)
                const uint32_t s = rs > rows ? 1u : 0u, rr = rs - s * (rows + 1u);
                float* sv = (s ? S.val[1] : S.val[0]) + rr * A.nx;
                const uint32_t yy = y0 + rr, gz = z + s + A.gz0;
                const bool row_face = yy == 0 || yy == A.ny - 1 || gz == 0 || gz == A.gnz - 1;  // domain faces in GLOBAL coordinates
                uint32_t bits[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t x = 128u * k + 32u * u + lane;
                    bits[u] = 0u;
                    if (x < A.nx) {
                        float val;
                        stage_point<MODE>(A, nd, x, row_face, craw[u], ca[u], cb[u], val, bits[u]);
                        if (MODE == M_BAND_RAW || !tma) sv[x] = val;  // the other modes interpolate the staged values as they are
                        if (ModeTraits<MODE>::kIds) (s ? S.idb[1] : S.idb[0])[rr * A.nx + x] = (unsigned char)(bits[u] & 3u);
                    }
                }
#pragma unroll
                for (int pl = 0; pl < NP; ++pl) {
                    uint4 wd;
                    wd.x = __ballot_sync(0xffffffffu, bits[0] & (4u << pl));
                    wd.y = __ballot_sync(0xffffffffu, bits[1] & (4u << pl));
                    wd.z = __ballot_sync(0xffffffffu, bits[2] & (4u << pl));
                    wd.w = __ballot_sync(0xffffffffu, bits[3] & (4u << pl));
                    if (lane == 0) *reinterpret_cast<uint4*>(S.plane + pl * plane_words + ((s * R1 + rr) * ppr + k) * 4u) = wd;
                }
                if (kAhead) {
                    k = nk; rs = nrs;
#pragma unroll
                    for (int u = 0; u < 4; ++u) { craw[u] = nraw[u]; ca[u] = na[u]; cb[u] = nb[u]; }
                } else {
                    k += kWarps;
                    while (k >= ppr) { k -= ppr; ++rs; }
                }
            }
#undef GCB_FETCH
        }
        __syncthreads();

        // ---- 2. classify: cube indices + counts.  Warp w owns a contiguous range of 128-cell steps (row r, chunk k) of the tile;
        //         lane l owns cells 128k + 4l .. + 3 of the step.
        const uint32_t nsteps = rows * cpr;
        const uint32_t spw = (nsteps + kWarps - 1) / kWarps;
        const uint32_t st_beg = min(warp * spw, nsteps), st_end = min(st_beg + spw, nsteps);
        const bool track = spw <= 32u && !A.st_verts;  // remember empty steps in a register mask
        const uint32_t r_beg = st_beg / cpr, k_beg = st_beg - r_beg * cpr;
        const uint32_t wj = lane >> 3, wsh = 4u * (lane & 7u);
        uint32_t emptymask = 0, my_verts = 0, my_act = 0;
        {
            uint32_t r = r_beg, k = k_beg;
            for (uint32_t st = st_beg; st < st_end; ++st) {
                const int left = (int)A.cx - (int)(128u * k + 4u * lane);
                const uint32_t vc = (uint32_t)max(0, min(4, left));          // valid cells of this lane
                const uint32_t rm = vc ? ((2u << vc) - 1u) : 0u;            // mask bits those cells look at
                uint32_t n[NP][4];
                bool lane_empty = true;
#pragma unroll
                for (int pl = 0; pl < NP; ++pl) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {  // q: slice << 1 | row offset
                        const uint32_t* w = S.plane + pl * plane_words + ((((uint32_t)q >> 1) * R1 + r + (q & 1)) * ppr + k) * 4u + wj;
                        n[pl][q] = __funnelshift_r(w[0], w[1], wsh) & 31u;
                    }
                    const uint32_t any = (n[pl][0] | n[pl][1] | n[pl][2] | n[pl][3]) & rm, all = n[pl][0] & n[pl][1] & n[pl][2] & n[pl][3] & rm;
                    lane_empty = lane_empty && (any == 0u || all == rm);
                }
                const bool empty = __all_sync(0xffffffffu, lane_empty);
                if (!(empty && track)) {
                    const uint32_t vmask = vc >= 4u ? 0xffffffffu : ((1u << (8u * vc)) - 1u);
                    uint32_t cw;
                    if (MODE == M_REGION) {
                        // one cube index per candidate test, then the reference's cascade per cell
                        constexpr int P1 = NP > 1 ? 1 : 0, P2 = NP > 2 ? 2 : 0;  // (only instantiated with NP == 3)
                        const uint32_t c0 = cubes4(n[0][0], n[0][1], n[0][2], n[0][3]) & vmask, c1 = cubes4(n[P1][0], n[P1][1], n[P1][2], n[P1][3]) & vmask,
                                       c2 = cubes4(n[P2][0], n[P2][1], n[P2][2], n[P2][3]) & vmask;
                        uint32_t clw = 0;
                        cw = 0;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const uint32_t k0 = (c0 >> (8 * u)) & 255u, k1 = (c1 >> (8 * u)) & 255u, k2 = (c2 >> (8 * u)) & 255u;
                            uint32_t cl = 0, cube = k0;
                            if (!(A.flags & F_SHOW_REGION) && cube == 0u) {
                                if (A.flags & F_SHOW_DOMAIN) { cube = k1; cl = 1; }
                                else { cube = k2; cl = 1; if (cube == 0u) { cube = k1; cl = 2; } }
                            }
                            cw |= cube << (8 * u);
                            clw |= cl << (8 * u);
                        }
                        *reinterpret_cast<uint32_t*>(S.cls + r * A.cstride + 128u * k + 4u * lane) = clw;
                    } else
                        cw = cubes4(n[0][0], n[0][1], n[0][2], n[0][3]) & vmask;
                    *reinterpret_cast<uint32_t*>(S.cube + r * A.cstride + 128u * k + 4u * lane) = cw;
                    const uint32_t e = (uint32_t)S.nv[cw & 255u] + S.nv[(cw >> 8) & 255u] + S.nv[(cw >> 16) & 255u] + S.nv[cw >> 24];
                    my_verts += e & 255u;
                    my_act += e >> 8;
                } else
                    emptymask |= 1u << (st - st_beg);
                if (++k == cpr) { k = 0; ++r; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            my_verts += __shfl_xor_sync(0xffffffffu, my_verts, o);
            my_act += __shfl_xor_sync(0xffffffffu, my_act, o);
        }
        if (lane == 0) { S.warp_tot[2 * warp] = my_verts; S.warp_tot[2 * warp + 1] = my_act; }
        __syncthreads();

        // ---- 3. decoupled look-back (warp 0)
        if (warp == 0) {
            uint32_t tv = 0, ta = 0;
            for (int w2 = 0; w2 < kWarps; ++w2) { tv += S.warp_tot[2 * w2]; ta += S.warp_tot[2 * w2 + 1]; }
            if (A.count_only) {
                if (lane == 0 && (tv | ta)) { atomicAdd(A.totals, (unsigned long long)ta); atomicAdd(A.totals + 1, (unsigned long long)tv); }
            } else {
                unsigned long long pa = 0, pv = 0;  // exclusive prefixes
                // A tile without active cells has nothing to place: it publishes a zero aggregate and does not wait for its own
                // prefix (sparse fields: most tiles, and the look-back wait was a quarter of their time).  Successors walk through
                // such entries like through any aggregate; every 64th tile still resolves and publishes an inclusive prefix, which
                // bounds the walk, and so does the last tile, which reports the totals.
                const bool no_prefix = ta == 0u && !A.st_verts && (tile & 63u) != 0u && tile != A.num_tiles - 1;
                if (no_prefix) {
                    if (lane == 0) { st_relaxed(A.status_a + tile, kFlagAgg); st_relaxed(A.status_v + tile, kFlagAgg); }
                } else if (tile > 0) {
                    if (lane == 0) { st_relaxed(A.status_a + tile, kFlagAgg | ta); st_relaxed(A.status_v + tile, kFlagAgg | tv); }
                    int64_t base = (int64_t)tile - 1;
                    for (;;) {
                        const int64_t idx = base - lane;
                        unsigned long long wa, wv;
                        if (idx >= 0) {
                            do {
                                wa = ld_relaxed(A.status_a + idx);
                                wv = ld_relaxed(A.status_v + idx);
                            } while ((wa >> 62) == 0ull || (wa >> 62) != (wv >> 62));
                        } else { wa = kFlagIncl; wv = kFlagIncl; }  // before tile 0: inclusive prefix 0
                        const uint32_t incl = __ballot_sync(0xffffffffu, (wa >> 62) == 2ull);
                        const uint32_t upto = incl ? (uint32_t)__ffs(incl) : 32u;  // lanes [0, upto) contribute
                        unsigned long long ca = (lane < upto) ? (wa & kValMask) : 0ull, cv = (lane < upto) ? (wv & kValMask) : 0ull;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            ca += __shfl_xor_sync(0xffffffffu, ca, o);
                            cv += __shfl_xor_sync(0xffffffffu, cv, o);
                        }
                        pa += ca; pv += cv;
                        if (incl) break;
                        base -= 32;
                    }
                }
                if (lane == 0 && !no_prefix) {
                    st_relaxed(A.status_a + tile, kFlagIncl | (pa + ta));
                    st_relaxed(A.status_v + tile, kFlagIncl | (pv + tv));
                    S.prefix[0] = pa; S.prefix[1] = pv;
                    if (tile == A.num_tiles - 1) { A.totals[0] = pa + ta; A.totals[1] = pv + tv; }
                }
            }
        }
        if (A.count_only) continue;  // next iteration's first __syncthreads orders smem reuse
        __syncthreads();

        // ---- 4. emit
        unsigned long long act_base = S.prefix[0], vert_base = S.prefix[1];
        for (uint32_t w2 = 0; w2 < warp; ++w2) { vert_base += S.warp_tot[2 * w2]; act_base += S.warp_tot[2 * w2 + 1]; }
        if (S.warp_tot[2 * warp] != 0 || A.st_verts) {
            uint32_t* q = S.queue + warp * kQueue;
            uint32_t head = 0, tail = 0, act_run = 0;  // triangles consumed / enqueued, active cells seen
            const size_t cell0 = (size_t)z * A.cx * A.cy + (size_t)y0 * A.cx;  // global id of tile cell 0 (local slab)
            uint32_t r = r_beg, k = k_beg;
            for (uint32_t st = st_beg; st < st_end; ++st, k = (k + 1 == cpr) ? 0u : k + 1u, r += (k == 0u)) {
                if ((emptymask >> (st - st_beg)) & 1u) continue;
                const uint32_t x0 = 128u * k + 4u * lane;
                const uint32_t cw = *reinterpret_cast<const uint32_t*>(S.cube + r * A.cstride + x0);
                uint32_t nt[4], ntl = 0, nactl = 0;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t e = S.nv[(cw >> (8 * u)) & 255u];
                    nt[u] = ((e & 255u) * 11u) >> 5;  // nv / 3 for nv in {0,3,..,15}
                    ntl += nt[u];
                    nactl += e >> 8;
                }
                // one packed inclusive scan: triangles in the low half, active cells in the high half (<= 640 / 128 per step)
                const uint32_t pk = ntl | (nactl << 16);
                uint32_t incl = pk;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t n2 = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= (uint32_t)o) incl += n2;
                }
                const uint32_t excl = incl - pk;
                const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
                const uint32_t step_tris = tot & 0xffffu, step_act = tot >> 16;
                if (step_act == 0u && !A.st_verts) continue;
                if (A.comp || A.st_verts) {
                    uint32_t rank = excl >> 16, tp = excl & 0xffffu;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const bool valid = x0 + u < A.cx;
                        const size_t gc = cell0 + (size_t)r * A.cx + x0 + u;
                        if (nt[u] > 0 && A.comp) A.comp[act_base + act_run + rank] = (uint32_t)gc + A.gz0 * A.cx * A.cy;
                        if (A.st_verts && valid) {
                            A.st_verts[gc] = 3u * nt[u];
                            A.st_occ[gc] = nt[u] > 0;
                            A.st_verts_scan[gc] = (uint32_t)(vert_base + 3ull * (tail + tp));
                            A.st_occ_scan[gc] = (uint32_t)(act_base + act_run + rank);
                        }
                        rank += nt[u] > 0;
                        tp += nt[u];
                    }
                }
                const uint32_t first = tail + (excl & 0xffffu), end = tail + step_tris;
                const uint32_t ebase = (r << 16) | x0;
                if (end - head <= (uint32_t)kQueue) {
                    // usual case: the whole step fits into the ring
                    uint32_t pos = first;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (nt[u] > 0) q[pos & (kQueue - 1)] = ebase + u;
                        if (nt[u] > 1) q[(pos + 1) & (kQueue - 1)] = (ebase + u) | (1u << 29);
                        if (__any_sync(0xffffffffu, nt[u] > 2)) {
                            if (nt[u] > 2) q[(pos + 2) & (kQueue - 1)] = (ebase + u) | (2u << 29);
                            if (nt[u] > 3) q[(pos + 3) & (kQueue - 1)] = (ebase + u) | (3u << 29);
                            if (nt[u] > 4) q[(pos + 4) & (kQueue - 1)] = (ebase + u) | (4u << 29);
                        }
                        pos += nt[u];
                    }
                    tail = end;
                    __syncwarp();
                    while (tail - head >= 32u) {
                        emit_triangle<MODE>(A, S, z, y0, q[(head + lane) & (kQueue - 1)], vert_base + 3ull * (head + lane));
                        head += 32u;
                    }
                    __syncwarp();
                } else {
                    // a step with more triangles than ring slots: enqueue position windows [lo, hi) and drain in between
                    uint32_t lo = tail;
                    for (;;) {
                        const uint32_t hi = (end - head <= (uint32_t)kQueue) ? end : head + kQueue;
                        uint32_t pos = first;
                        for (int u = 0; u < 4; ++u)
                            for (uint32_t jj = 0; jj < nt[u]; ++jj, ++pos)
                                if (pos - lo < hi - lo) q[pos & (kQueue - 1)] = (ebase + u) | (jj << 29);
                        tail = hi;
                        __syncwarp();
                        while (tail - head >= 32u) {
                            emit_triangle<MODE>(A, S, z, y0, q[(head + lane) & (kQueue - 1)], vert_base + 3ull * (head + lane));
                            head += 32u;
                        }
                        __syncwarp();
                        if (hi == end) break;
                        lo = hi;
                    }
                }
                act_run += step_act;
            }
            if (lane < tail - head) emit_triangle<MODE>(A, S, z, y0, q[(head + lane) & (kQueue - 1)], vert_base + 3ull * (head + lane));
        }
    }
}

// ---------------------------------------------------------------- host side
template <int MODE, bool TMA>
static cudaError_t launch_one(const McArgs& a, int grid, size_t smem, cudaStream_t st) {
    mc_fused_kernel<MODE, TMA><<<grid, kThreads, smem, st>>>(a);  // cudaFuncAttributeMaxDynamicSharedMemorySize: set in occupancy_one()
    return cudaGetLastError();
}
template <int MODE>
static cudaError_t launch_mode(const McArgs& a, int grid, size_t smem, cudaStream_t st) {
    return a.use_tma ? launch_one<MODE, true>(a, grid, smem, st) : launch_one<MODE, false>(a, grid, smem, st);
}

template <int MODE, bool TMA>
static int occupancy_one(size_t smem, size_t attr_smem) {
    int n = 0;
    if (cudaFuncSetAttribute(mc_fused_kernel<MODE, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attr_smem) != cudaSuccess) return 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, mc_fused_kernel<MODE, TMA>, kThreads, smem);
    return n;
}
template <int MODE>
static int occupancy(size_t smem, size_t attr_smem, bool tma) {
    return tma ? occupancy_one<MODE, true>(smem, attr_smem) : occupancy_one<MODE, false>(smem, attr_smem);
}

int launch_extract(Ctx* c, McArgs& a, unsigned long long* active, unsigned long long* verts, unsigned long long* h_totals_async) {
    if (a.nx < 2 || a.ny < 2 || a.nz < 2) {
        if (h_totals_async) { h_totals_async[0] = 0; h_totals_async[1] = 0; }
        else { *active = 0; *verts = 0; }
        return 0;
    }
    a.cx = a.nx - 1; a.cy = a.ny - 1; a.cz = a.nz - 1;
    // tile height: staged rows must fit in shared memory
    if (a.nx > 65535u) return fail_msg(c, "grid row too wide (nx > 65535)");
    a.ppr = (a.nx + 127u) / 128u;
    a.cpr = (a.cx + 127u) / 128u;
    a.cstride = 128u * a.cpr;
    // ~4096 cells per tile; large grids in the three-CTA modes take twice that (measured on the 768x384x384 density: 8 rows per
    // tile instead of 5 at the same occupancy, 1.29 -> 1.09 ms -- the per-tile latencies are what a sparse surface pays for)
    uint32_t tile_cells = (a.mode != M_BAND_RAW && (unsigned long long)a.cx * a.cy * a.cz >= (1ull << 25)) ? 8192u : 4096u;
    if (const char* e = getenv("GCB_MC_TILE_CELLS")) tile_cells = (uint32_t)atoi(e);  // A/B measurements
    uint32_t R = tile_cells / a.cx;
    if (R < 1) R = 1;
    if (R > a.cy) R = a.cy;
    const bool ids = a.mode == M_LATTICE_ONE || a.mode == M_LATTICE;
    const size_t planes = a.mode == M_REGION ? 3 : 1;
    const size_t fixed_bytes = 256 * 8 + 16 + 8 + 8 + 32 * 8 + kWarps * kQueue * 4 + kWarps * 8 + 512 + 64 /*slack*/;
    auto smem_for = [&](uint32_t r) {
        const size_t stride = (((size_t)(r + 1) * a.nx) + 15) & ~(size_t)15;
        return stride * 4 * 2 + (ids ? stride * 2 : 0) + planes * 2 * (r + 1) * a.ppr * 16 + 16 + (size_t)r * a.cstride * (a.mode == M_REGION ? 2 : 1) + fixed_bytes;
    };
    // 228 KB shared memory per SM, 1 KB reserved per CTA: 56 KB keeps four CTAs on an SM (M_BAND_RAW's register budget; 6-row tiles
    // for 512-wide rows), 75 KB three (the other modes).  Rows of 2048 points leave one row per tile
    // at 56 KB -- every staged row is then shared by a single cell row -- so very wide grids trade the fourth CTA for taller
    // tiles (2048-wide slab: 1 row per tile at 56 KB 5.52 ms, 2 rows at 75 KB 5.05 ms; 1024-wide: 3 rows at 56 KB 2.49 ms, 5 rows
    // at 75 KB 2.58 ms; equal at 512).
    const uint32_t R0 = R;
    size_t cap = (a.mode == M_BAND_RAW ? 56 : 75) * 1024;
    while (R > 1 && smem_for(R) > cap) --R;
    if (a.mode == M_BAND_RAW && R < 3 && R < R0) { cap = 75 * 1024; R = R0; }
    if (const char* e = getenv("GCB_MC_SMEM_CAP_KB")) { cap = (size_t)atoi(e) * 1024; R = R0; }  // A/B measurements of tile height vs CTAs per SM
    while (R > 1 && smem_for(R) > cap) --R;
    const size_t smem = smem_for(R);
    if (smem > 227 * 1024) return fail_msg(c, "grid row too wide for one shared-memory tile (nx too large)");
    if (R > 8191u) return fail_msg(c, "tile too tall");
    a.rows_per_tile = R;
    a.tiles_per_slice = (a.cy + R - 1) / R;
    const unsigned long long nt = (unsigned long long)a.tiles_per_slice * a.cz;
    if (nt >= 0xffffffffull) return fail_msg(c, "too many tiles");
    a.num_tiles = (uint32_t)nt;
    a.prow_stride = (uint32_t)((((size_t)(R + 1) * a.nx) + 15) & ~(size_t)15);
    {
        float thr = (float)0.0005;
        if ((double)thr < 0.0005) thr = nextafterf(thr, 1.0f);
        a.snap_thr = thr;
    }
    // TMA bulk copies need 16-byte aligned global addresses and sizes
    a.use_tma = !(c->options & GCB_OPT_NO_TMA) && a.f0 && (a.nx % 4 == 0) && (((uintptr_t)a.f0 & 15) == 0) && (((uintptr_t)a.f0_top & 15) == 0);
    a.top_begin = (unsigned long long)(a.nz - 1) * a.nx * a.ny;
    if ((a.f0_top || a.f1_top || a.gp_top) && (a.mode == M_REGION || a.mode == M_BAND_RAW || (a.flags & F_DISP) || a.f2))
        return fail_msg(c, "peer-held halo layer: supported for the latticeone / CSG / topo inputs f0, f1 and grid_points only");

    // one scratch allocation, cleared by ONE memset per launch: [totals (2 words) | tile counter | pad | status_a | status_v]
    const size_t need = 4 + 2 * (size_t)a.num_tiles;
    if (c->status_cap < need) {
        if (c->d_status) cudaFree(c->d_status);
        c->status_cap = need + 4096;
        GCB_CHECK(c, cudaMalloc(&c->d_status, c->status_cap * sizeof(unsigned long long)));
    }
    a.totals = c->d_status;
    a.tile_counter = (uint32_t*)(c->d_status + 2);
    a.status_a = c->d_status + 4;
    a.status_v = c->d_status + 4 + a.num_tiles;
    GCB_CHECK(c, cudaMemsetAsync(c->d_status, 0, (a.count_only ? 4 : need) * sizeof(unsigned long long), c->stream));

    // occupancy query + shared-memory attribute once per (device, mode, stage path, tile size): both are host-side driver calls
    // that a small grid would otherwise pay on every launch.  The attribute belongs to the kernel FUNCTION (per device), not to
    // a gcb context, so the cache is process-wide, keyed by device, and guarded by a mutex: contexts on several host threads may
    // launch concurrently.  The attribute only ever grows (a launch may request less dynamic shared memory than the function's
    // maximum, never more), so a context that still uses a smaller tile is never invalidated by another one's larger tile.
    if (a.mode < 0 || a.mode > M_REGION) return fail_msg(c, "bad mode");
    struct OccEntry { size_t attr_smem = 0; size_t smem = 0; int occ = 0; };
    static std::mutex occ_mutex;
    static std::map<std::tuple<int, int, int>, OccEntry> occ_cache;
    int occ = 1;
    {
        std::lock_guard<std::mutex> lock(occ_mutex);
        OccEntry& oc = occ_cache[std::make_tuple(c->device, a.mode, a.use_tma ? 1 : 0)];
        if (oc.smem == smem && oc.occ > 0 && oc.attr_smem >= smem) occ = oc.occ;
        else {
            const size_t attr = std::max(oc.attr_smem, smem);
            switch (a.mode) {
            case M_LATTICE_ONE: occ = occupancy<M_LATTICE_ONE>(smem, attr, a.use_tma); break;
            case M_LATTICE: occ = occupancy<M_LATTICE>(smem, attr, a.use_tma); break;
            case M_CSG: occ = occupancy<M_CSG>(smem, attr, a.use_tma); break;
            case M_TOPO: occ = occupancy<M_TOPO>(smem, attr, a.use_tma); break;
            case M_BAND_RAW: occ = occupancy<M_BAND_RAW>(smem, attr, a.use_tma); break;
            case M_REGION: occ = occupancy<M_REGION>(smem, attr, a.use_tma); break;
            default: return fail_msg(c, "bad mode");
            }
            oc.smem = smem; oc.occ = occ; oc.attr_smem = attr;
        }
    }
    if (occ < 1) return fail_msg(c, "extraction kernel does not fit on an SM");
    // persistent grid: every CTA resident (required by the look-back's forward progress)
    long long grid = (long long)occ * c->num_sms;
    if (grid > (long long)a.num_tiles) grid = a.num_tiles;

    if (c->timing) cudaEventRecord(c->ev[0], c->stream);
    cudaError_t e;
    switch (a.mode) {
    case M_LATTICE_ONE: e = launch_mode<M_LATTICE_ONE>(a, (int)grid, smem, c->stream); break;
    case M_LATTICE: e = launch_mode<M_LATTICE>(a, (int)grid, smem, c->stream); break;
    case M_CSG: e = launch_mode<M_CSG>(a, (int)grid, smem, c->stream); break;
    case M_TOPO: e = launch_mode<M_TOPO>(a, (int)grid, smem, c->stream); break;
    case M_REGION: e = launch_mode<M_REGION>(a, (int)grid, smem, c->stream); break;
    default: e = launch_mode<M_BAND_RAW>(a, (int)grid, smem, c->stream); break;
    }
    if (e != cudaSuccess) return fail(c, "mc_fused_kernel launch", e);
    c->launches++;
    if (c->timing) { cudaEventRecord(c->ev[1], c->stream); c->extract_timed = true; }
    unsigned long long* dst = h_totals_async ? h_totals_async : c->h_totals;
    GCB_CHECK(c, cudaMemcpyAsync(dst, a.totals, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    if (h_totals_async) return 0;  // enqueue only: the caller synchronises and reads {active, vertices} from its pinned slot
    GCB_CHECK(c, cudaStreamSynchronize(c->stream));
    *active = c->h_totals[0];
    *verts = c->h_totals[0] ? c->h_totals[1] : 0;  // early-out of Isosurface.cu:83-87
    return 0;
}

} // namespace gcb
