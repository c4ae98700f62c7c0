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
//                 global -> shared with TMA bulk copies (cp.async.bulk + mbarrier); each staged
//                 point is turned into {value, inside-bit, id-class} (mode specific);
//   2. classify : cube index per cell from the staged bits, vertex count from the table (smem
//                 copy of the __constant__ table: per-lane indices diverge), warp/block totals;
//   3. look-back: single-pass decoupled look-back over tiles publishes {active, vertex} prefixes,
//                 so offsets are exactly the reference's exclusive scans;
//   4. emit     : active cells append their triangles to a per-warp queue; the warp drains the
//                 queue 32 triangles at a time (one triangle per lane, vertices interpolated from
//                 the STAGED field) and writes float4 pos/norm at the reference's indices.
// The field is read from HBM once; no per-cell scratch arrays are written (the reference moves
// 36-56 B/cell through them).
#include "common.cuh"

#include <cmath>

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
    unsigned char* bit[2];
    unsigned char* cube;
    unsigned char* cls;   // M_REGION: which cascade level produced the cube index (0: vol_topo, 1: second test, 2: third test)
    unsigned long long* tri;
    unsigned char* nv;
    uint32_t* queue;      // kWarps * kQueue
    uint32_t* warp_tot;   // kWarps * 2
    unsigned long long* prefix;  // [0]=active prefix, [1]=vertex prefix (exclusive, for this tile)
    uint32_t* tile_id;
    uint64_t* mbar;
    unsigned char* edge;  // [0..11] lattice edge order, [16..27] owner edge order: corner bits ca | cb << 3
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
__device__ __forceinline__ float3 interp_band(float thr, float l1, float l2, float3 p0, float3 p1, float f0, float f1, uint32_t id0, uint32_t id1) {
    // Branch-free restatement (lanes of a warp hold unrelated edges, so every if/else of the reference would serialise):
    // all candidates are computed, the reference's decision tree only selects.
    const bool crossing = (id0 == 1u && id1 == 0u) || (id0 == 0u && id1 == 1u);
    const bool sw = f1 < f0;
    const float lo = sw ? f1 : f0, hi = sw ? f0 : f1;
    const float3 plo = sw ? p1 : p0, phi = sw ? p0 : p1;
    const bool c1 = (hi >= l1) && (lo <= l1);
    const bool c2 = !c1 && (hi >= l2) && (lo <= l2);
    const float lv = c1 ? l1 : l2;
    const float dn = __fsub_rn(lv, lo), dd = __fsub_rn(hi, lo);
    const bool s_lo = snap(dn, thr) || (!snap(__fsub_rn(lv, hi), thr) && snap(dd, thr));  // -> p0 (1st or 3rd test)
    const bool s_hi = !snap(dn, thr) && snap(__fsub_rn(lv, hi), thr);                        // -> p1 (2nd test)
    float t = __fdiv_rn(dn, dd);
    if (!(c1 || c2)) t = (hi == lo && plo.z == 0.0f) ? 1.f : 0.f;  // reference: t = 1 / t = 0 / (unset -> 0 here)
    if (!crossing) t = 0.f;
    // without a crossing the reference lerps the UNSWAPPED endpoints with t = 0, i.e. returns p0 + 0*(p1-p0)
    const float3 a = crossing ? plo : p0, bb = crossing ? phi : p1;
    float3 r = lerp3(a, bb, t);
    if (crossing && (c1 || c2)) {
        if (s_lo) r = plo;
        else if (s_hi) r = phi;
    }
    return r;
}
// The same for M_BAND_RAW, where the mask is built in-kernel from the band test: an edge named by the triangle table joins
// a point with mask 0 to a point with mask 1 (the cube bits ARE "mask < iso" and the mask only takes the values 0 and 1), so
// the reference's id test is true for every edge that reaches this function and the ids need not be staged or looked at.
__device__ __forceinline__ float3 interp_band_crossing(float thr, float l1, float l2, float3 p0, float3 p1, float f0, float f1) {
    const bool sw = f1 < f0;
    const float lo = sw ? f1 : f0, hi = sw ? f0 : f1;
    const float3 plo = sw ? p1 : p0, phi = sw ? p0 : p1;
    const bool c1 = (hi >= l1) && (lo <= l1);
    const bool c2 = !c1 && (hi >= l2) && (lo <= l2);
    const float lv = c1 ? l1 : l2;
    const float dn = __fsub_rn(lv, lo), dd = __fsub_rn(hi, lo);
    const bool s_lo = snap(dn, thr) || (!snap(__fsub_rn(lv, hi), thr) && snap(dd, thr));
    const bool s_hi = !snap(dn, thr) && snap(__fsub_rn(lv, hi), thr);
    float t = __fdiv_rn(dn, dd);
    if (!(c1 || c2)) t = (hi == lo && plo.z == 0.0f) ? 1.f : 0.f;
    float3 r = lerp3(plo, phi, t);
    if (c1 || c2) {
        if (s_lo) r = plo;
        else if (s_hi) r = phi;
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
struct UniformDiv { float d, y; bool ok; };
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

template <int MODE>
__device__ __forceinline__ void stage_point(const McArgs& A, const UniformDiv& nd, size_t gi, uint32_t x, bool row_face, float raw, float& val, uint32_t& bits) {
    if (MODE == M_LATTICE_ONE || MODE == M_LATTICE) {
        const float m = __ldg(A.f1 + gi);  // mask `vol`; classifyVoxel_new :3232-3239
        uint32_t id = (m == 1.f) ? 1u : (m == 0.f) ? 0u : (m == 2.f) ? 2u : 3u;
        bits = id | ((m < A.iso) ? 4u : 0u);
        val = raw;  // k `vol_one`
    } else if (MODE == M_BAND_RAW) {
        // device_bufferfour (Gratings.cu:1089-1134) fused; domain faces use GLOBAL coordinates
        float k = div_by_uniform(__fsub_rn(raw, A.na), nd);
        float m;
        if (row_face || x == 0 || x == A.nx - 1) { m = 0.0f; k = 0.0f; }
        else m = ((k >= A.iso1) && (k <= A.iso2)) ? 1.0f : 0.0f;
        bits = (m == 1.f ? 1u : 0u) | ((m < A.iso) ? 4u : 0u);
        val = k;
    } else if (MODE == M_REGION) {
        // classifyVoxel_region_kernel :1163-1290: three candidate inside tests per point, the cascade is resolved per cell
        const float iso = A.iso;
        const float ft = A.gp2 ? (float)A.gp2[gi].val : 0.f;  // sampleVolume_2 :98-108
        const float fx = A.gp ? (float)A.gp[gi].val : 0.f;
        bits = ((ft < iso) ? 4u : 0u) | ((fx < iso) ? 8u : 0u) | (((fx < iso) & (raw < iso)) ? 16u : 0u);
        val = raw;  // primitive_dynamic
    } else if (MODE == M_TOPO) {
        // classifyVoxel_kernel_topo :1492-1499
        const float fx = A.gp ? (float)A.gp[gi].val : 0.f;
        bits = ((fx < A.iso1) | (raw >= A.iso)) ? 4u : 0u;
        val = raw;
    } else {  // M_CSG  classifyVoxel :922-1052
        const float iso = A.iso;
        const float fx = A.gp ? (float)A.gp[gi].val : 0.f;
        const float dy = raw;
        const bool fixed = A.flags & F_FIXED, dyn = A.flags & F_DYNAMIC;
        float la = 0.f;
        if ((fixed || dyn) && A.f1) la = __ldg(A.f1 + gi);
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
template <int MODE>
__device__ __forceinline__ void emit_triangle(const McArgs& A, const Smem& S, uint32_t z, uint32_t y0, uint32_t c, uint32_t j,
                                              unsigned long long vidx) {
    const uint32_t r = (uint32_t)(((unsigned long long)c * A.magic_cx_mul) >> A.magic_cx_shift), x = c - r * A.cx;  // c / cx, exact for c < 2^28
    const uint32_t cube = S.cube[c];
    const unsigned long long tri = S.tri[cube];
    const uint32_t y = y0 + r;
    // MarchingCubes_kernel.cu:1888-1890 : (uint -> float) - center, times voxel
    const float3 p = make_float3(__fmul_rn(__fsub_rn((float)x, A.center.x), A.voxel.x), __fmul_rn(__fsub_rn((float)y, A.center.y), A.voxel.y),
                                 __fmul_rn(__fsub_rn((float)(z + A.gz0), A.center.z), A.voxel.z));  // global z under slab sharding

    const float3 pmax = make_float3(__fadd_rn(p.x, A.voxel.x), __fadd_rn(p.y, A.voxel.y), __fadd_rn(p.z, A.voxel.z));
    const uint32_t sbase = r * A.nx + x;
    float3 v[3];
    float w[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const uint32_t e = (uint32_t)(tri >> (4 * (3 * j + k))) & 15u;
        const bool own = (MODE == M_TOPO) || (MODE == M_REGION) || (MODE == M_CSG && !(A.flags & F_FIXED));
        const uint32_t ab = S.edge[(own ? 16u : 0u) + e];
        const uint32_t ca = ab & 7u, cb = ab >> 3;
        // corner positions: v[0] = p, v[i] = p + (voxel or 0) per component (:1892-1900).  p + 0.0f == p bit for bit for every
        // p this kernel can produce (x - center is never -0, voxel sizes are positive), so each component is a select.
        const float3 pa = make_float3((ca & 1u) ? pmax.x : p.x, (ca & 2u) ? pmax.y : p.y, (ca & 4u) ? pmax.z : p.z);
        const float3 pb = make_float3((cb & 1u) ? pmax.x : p.x, (cb & 2u) ? pmax.y : p.y, (cb & 4u) ? pmax.z : p.z);
        const uint32_t sa = sbase + ((ca & 2u) ? A.nx : 0u) + (ca & 1u), sb = sbase + ((cb & 2u) ? A.nx : 0u) + (cb & 1u);
        const uint32_t za = (ca >> 2) & 1u, zb = (cb >> 2) & 1u;
        const float fa = (za ? S.val[1] : S.val[0])[sa], fb = (zb ? S.val[1] : S.val[0])[sb];
        w[k] = 0.f;
        if (MODE == M_REGION) {
            // generateTriangles_region_kernel :2391-2490: stored crossing parameter of the edge's owning point, from vol_topo
            // (first test fired) or primitive_fixed; make_region's second test blends it with the dynamic field's crossing
            const size_t ga = ((size_t)(z + za) * A.ny + y + ((ca >> 1) & 1u)) * A.nx + x + (ca & 1u);
            const uint32_t cl = S.cls[c];
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
        } else if (MODE == M_BAND_RAW) {
            v[k] = interp_band_crossing(A.snap_thr, A.iso1, A.iso2, pa, pb, fa, fb);
        } else if (MODE == M_LATTICE_ONE) {
            v[k] = interp_band(A.snap_thr, A.iso1, A.iso2, pa, pb, fa, fb, ((za ? S.bit[1] : S.bit[0])[sa] & 3u), ((zb ? S.bit[1] : S.bit[0])[sb] & 3u));
        } else if (MODE == M_LATTICE) {
            const uint32_t ida = ((za ? S.bit[1] : S.bit[0])[sa] & 3u), idb = ((zb ? S.bit[1] : S.bit[0])[sb] & 3u);
            if ((ida == 2u && idb == 0u) || (idb == 2u && ida == 0u)) {
                // ids {2,0}: first block of vertexInterp3_new does not fire, second one does
                const size_t ga = ((size_t)(z + za) * A.ny + y + ((ca >> 1) & 1u)) * A.nx + x + (ca & 1u);
                const size_t gb = ((size_t)(z + zb) * A.ny + y + ((cb >> 1) & 1u)) * A.nx + x + (cb & 1u);
                float t = 0.f;
                float3 out;
                float3 q0 = pa, q1 = pb;
                if (interp_band_two(A.snap_thr, A.iso1b, A.iso2b, q0, q1, __ldg(A.f2 + ga), __ldg(A.f2 + gb), t, out)) v[k] = out;
                else v[k] = lerp3(q0, q1, t);
            } else v[k] = interp_band(A.snap_thr, A.iso1, A.iso2, pa, pb, fa, fb, ida, idb);
        } else {
            const size_t ga = ((size_t)(z + za) * A.ny + y + ((ca >> 1) & 1u)) * A.nx + x + (ca & 1u);
            const size_t gb = ((size_t)(z + zb) * A.ny + y + ((cb >> 1) & 1u)) * A.nx + x + (cb & 1u);
            float et = 0.f;
            if (own && A.gp) {
                const GridPoint g = A.gp[ga];
                const uint32_t ax = edge_axis(e);
                et = ax == 0 ? g.t_x : ax == 1 ? g.t_y : g.t_z;
            }
            float t;
            float3 qa = pa, qb = pb;
            if (MODE == M_TOPO) {
                t = t_analysis(A.iso, fa, fb, et);
                w[k] = A.f1 ? __ldg(A.f1 + ga) : 0.f;  // *field_val = r0 (:1797)
                if (A.flags & F_DISP) {
                    const float4 d0 = A.disp[ga], d1 = A.disp[gb];
                    qa = make_float3(d0.x, d0.y, d0.z);
                    qb = make_float3(d1.x, d1.y, d1.z);
                }
            } else if (A.flags & F_MAKE_REGION) t = et;
            else if (A.flags & F_FIXED) t = t_fixed(A.iso, A.iso1, A.iso2, fa, fb, __ldg(A.f1 + ga), __ldg(A.f1 + gb));
            else if (A.flags & F_DYNAMIC) t = t_primitive_one(A.iso1, A.iso2, __ldg(A.f1 + ga), __ldg(A.f1 + gb), et);
            else t = t_primitive(A.iso, fa, fb, et);
            v[k] = lerp3(qa, qb, t);
        }
    }
    float3 n;
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
    const unsigned long long limit = (unsigned long long)((unsigned int)A.max_verts - 3u);  // uint wrap as :2181
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
template <int MODE>
__global__ void __launch_bounds__(kThreads) mc_fused_kernel(const McArgs A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem S;
    {
        unsigned char* p = smem_raw;
        S.val[0] = (float*)p; p += (size_t)A.prow_stride * 4;
        S.val[1] = (float*)p; p += (size_t)A.prow_stride * 4;
        S.tri = (unsigned long long*)p; p += 256 * 8;
        S.prefix = (unsigned long long*)p; p += 16;
        S.mbar = (uint64_t*)p; p += 8;
        S.tile_id = (uint32_t*)p; p += 8;
        S.edge = p; p += 32;
        S.queue = (uint32_t*)p; p += kWarps * kQueue * 4;
        S.warp_tot = (uint32_t*)p; p += kWarps * 2 * 4;
        S.nv = p; p += 256;
        S.bit[0] = p; p += A.prow_stride;
        S.bit[1] = p; p += A.prow_stride;
        S.cube = p; p += (size_t)A.rows_per_tile * A.cx;
        S.cls = p;  // only backed by shared memory in M_REGION launches
    }
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

    for (uint32_t i = tid; i < 256; i += kThreads) {
        const unsigned long long t = c_tri_packed[i];
        S.tri[i] = t;
        // number of used nibbles = numVertsTable[i]
        unsigned long long u = ~t;                      // used nibble != 0xF  <=> ~nibble != 0
        u = (u | (u >> 1) | (u >> 2) | (u >> 3)) & 0x1111111111111111ull;
        S.nv[i] = (unsigned char)__popcll(u);
    }
    if (tid < 12) {
        const uint32_t l = edge_lat(tid), o = edge_own(tid);
        S.edge[tid] = (unsigned char)(corner_bits(l & 15u) | (corner_bits(l >> 4) << 3));
        S.edge[16 + tid] = (unsigned char)(corner_bits(o & 15u) | (corner_bits(o >> 4) << 3));
    }
    if (tid == 0) { mbar_init(S.mbar, 1); fence_mbar_init(); }
    __syncthreads();

    uint32_t parity = 0;
    const uint32_t slice_pts = A.nx * A.ny;
    const UniformDiv nd = make_uniform_div(__fsub_rn(A.nb, A.na));  // M_BAND_RAW normalisation range

    for (;;) {
        if (tid == 0) *S.tile_id = atomicAdd(A.tile_counter, 1u);
        __syncthreads();  // also orders the previous tile's smem reads before this tile's writes
        const uint32_t tile = *S.tile_id;
        if (tile >= A.num_tiles) break;
        const uint32_t z = tile / A.tiles_per_slice, ty = tile - z * A.tiles_per_slice;
        const uint32_t y0 = ty * A.rows_per_tile;
        const uint32_t rows = min(A.rows_per_tile, A.cy - y0);
        const uint32_t npts = (rows + 1) * A.nx;   // staged points per slice
        const uint32_t ncell = rows * A.cx;
        const size_t g0 = (size_t)z * slice_pts + (size_t)y0 * A.nx;  // first staged point, slice z

        // ---- 1. stage-in
        if (A.use_tma && A.f0) {
            if (tid == 0) {
                fence_proxy_async();  // generic-proxy accesses of the previous tile precede the async writes
                mbar_expect_tx(S.mbar, 2u * npts * 4u);
                tma_bulk_g2s(S.val[0], A.f0 + g0, npts * 4u, S.mbar);
                tma_bulk_g2s(S.val[1], A.f0 + g0 + slice_pts, npts * 4u, S.mbar);
            }
            mbar_wait(S.mbar, parity);
            parity ^= 1u;
        }
#pragma unroll 1
        for (uint32_t s = 0; s < 2; ++s) {
            float* sv = s ? S.val[1] : S.val[0];
            unsigned char* sb = s ? S.bit[1] : S.bit[0];
            const size_t gs = g0 + (size_t)s * slice_pts;
            for (uint32_t rr = warp; rr <= rows; rr += kWarps) {
                const uint32_t yy = y0 + rr, gz = z + s + A.gz0;
                const bool row_face = yy == 0 || yy == A.ny - 1 || gz == 0 || gz == A.gnz - 1;  // domain faces in GLOBAL coordinates
                if (A.use_tma && A.f0) {
                    // TMA path: nx % 4 == 0 and rows are 16-byte aligned in shared memory -> four points per lane and iteration
                    for (uint32_t x = lane * 4; x < A.nx; x += 128) {
                        const uint32_t pnt = rr * A.nx + x;
                        float4 v4 = *reinterpret_cast<const float4*>(sv + pnt);
                        float vv[4] = {v4.x, v4.y, v4.z, v4.w};
                        uint32_t packed = 0;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            float val;
                            uint32_t bits;
                            stage_point<MODE>(A, nd, gs + pnt + u, x + u, row_face, vv[u], val, bits);
                            vv[u] = val;
                            packed |= bits << (8 * u);
                        }
                        *reinterpret_cast<float4*>(sv + pnt) = make_float4(vv[0], vv[1], vv[2], vv[3]);
                        *reinterpret_cast<uint32_t*>(sb + pnt) = packed;
                    }
                    continue;
                }
                for (uint32_t x = lane; x < A.nx; x += 32) {
                    const uint32_t pnt = rr * A.nx + x;
                    float raw;
                    if (A.use_tma && A.f0) raw = sv[pnt];
                    else raw = A.f0 ? __ldg(A.f0 + gs + pnt) : 0.f;
                    float val;
                    uint32_t bits;
                    stage_point<MODE>(A, nd, gs + pnt, x, row_face, raw, val, bits);
                    sv[pnt] = val;
                    sb[pnt] = (unsigned char)bits;
                }
            }
        }
        __syncthreads();

        // ---- 2. classify: cube index + counts.  Warp w owns a contiguous cell range of the tile.
        const uint32_t seg = (((ncell + kWarps - 1) / kWarps) + 31u) & ~31u;
        const uint32_t cbeg = min(warp * seg, ncell), cend = min(cbeg + seg, ncell);
        uint32_t my_verts = 0, my_act = 0;
        {
            uint32_t c = cbeg + lane;
            uint32_t r = c / A.cx, x = c - r * A.cx;
            for (; c < cend; c += 32) {
                const uint32_t pi = r * A.nx + x;
                const unsigned char* b0 = S.bit[0] + pi;
                const unsigned char* b1 = S.bit[1] + pi;
                uint32_t cube;
                if (MODE == M_REGION) {
                    // the eight corner bytes, then one cube index per candidate test (bits 2, 3, 4) and the reference's cascade
                    const uint32_t cb[8] = {b0[0], b0[1], b0[A.nx + 1], b0[A.nx], b1[0], b1[1], b1[A.nx + 1], b1[A.nx]};
                    uint32_t k0 = 0, k1 = 0, k2 = 0;
#pragma unroll
                    for (int q = 0; q < 8; ++q) { k0 |= ((cb[q] >> 2) & 1u) << q; k1 |= ((cb[q] >> 3) & 1u) << q; k2 |= ((cb[q] >> 4) & 1u) << q; }
                    uint32_t cl = 0;
                    cube = k0;
                    if (!(A.flags & F_SHOW_REGION) && cube == 0u) {
                        if (A.flags & F_SHOW_DOMAIN) { cube = k1; cl = 1; }
                        else { cube = k2; cl = 1; if (cube == 0u) { cube = k1; cl = 2; } }
                    }
                    S.cls[c] = (unsigned char)cl;
                } else
                    cube = ((b0[0] >> 2) & 1u) | (((b0[1] >> 2) & 1u) << 1) | (((b0[A.nx + 1] >> 2) & 1u) << 2) | (((b0[A.nx] >> 2) & 1u) << 3) |
                           (((b1[0] >> 2) & 1u) << 4) | (((b1[1] >> 2) & 1u) << 5) | (((b1[A.nx + 1] >> 2) & 1u) << 6) | (((b1[A.nx] >> 2) & 1u) << 7);
                S.cube[c] = (unsigned char)cube;
                const uint32_t nv = S.nv[cube];
                my_verts += nv;
                my_act += nv > 0;
                x += 32;
                while (x >= A.cx) { x -= A.cx; ++r; }
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
                if (tile > 0) {
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
                if (lane == 0) {
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
            for (uint32_t cb = cbeg; cb < cend; cb += 32) {
                const uint32_t c = cb + lane;
                const bool valid = c < cend;
                const uint32_t nv = valid ? S.nv[S.cube[c]] : 0u;
                const uint32_t nt = (nv * 11u) >> 5;  // nv / 3 for nv in {0,3,..,15}
                const uint32_t amask = __ballot_sync(0xffffffffu, nv > 0);
                if (amask == 0u && !A.st_verts) continue;  // nothing to scan, enqueue or drain in this 32-cell step
                uint32_t incl = nt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t n2 = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= (uint32_t)o) incl += n2;
                }
                const uint32_t excl = incl - nt;
                const uint32_t step_tris = __shfl_sync(0xffffffffu, incl, 31);
                const uint32_t rank = __popc(amask & ((1u << lane) - 1u));
                if (nv > 0 && A.comp) A.comp[act_base + act_run + rank] = (uint32_t)(cell0 + c) + A.gz0 * A.cx * A.cy;
                if (A.st_verts && valid) {
                    const size_t gc = cell0 + c;
                    A.st_verts[gc] = nv;
                    A.st_occ[gc] = nv > 0;
                    A.st_verts_scan[gc] = (uint32_t)(vert_base + 3ull * (tail + excl));
                    A.st_occ_scan[gc] = (uint32_t)(act_base + act_run + rank);
                }
                for (uint32_t jj = 0; __any_sync(0xffffffffu, jj < nt); ++jj)  // at most 5 rounds, warp-uniform trip count
                    if (jj < nt) q[(tail + excl + jj) & (kQueue - 1)] = (jj << 28) | c;
                tail += step_tris;
                act_run += __popc(amask);
                __syncwarp();
                while (tail - head >= 32u) {
                    const uint32_t ent = q[(head + lane) & (kQueue - 1)];
                    emit_triangle<MODE>(A, S, z, y0, ent & 0x0fffffffu, ent >> 28, vert_base + 3ull * (head + lane));
                    head += 32u;
                }
                __syncwarp();
            }
            if (lane < tail - head) {
                const uint32_t ent = q[(head + lane) & (kQueue - 1)];
                emit_triangle<MODE>(A, S, z, y0, ent & 0x0fffffffu, ent >> 28, vert_base + 3ull * (head + lane));
            }
        }
    }
}

// ---------------------------------------------------------------- host side
template <int MODE>
static cudaError_t launch_mode(const McArgs& a, int grid, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(mc_fused_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    mc_fused_kernel<MODE><<<grid, kThreads, smem, st>>>(a);
    return cudaGetLastError();
}

template <int MODE>
static int occupancy(size_t smem) {
    int n = 0;
    cudaFuncSetAttribute(mc_fused_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, mc_fused_kernel<MODE>, kThreads, smem);
    return n;
}

int launch_extract(Ctx* c, McArgs& a, unsigned long long* active, unsigned long long* verts) {
    if (a.nx < 2 || a.ny < 2 || a.nz < 2) { *active = 0; *verts = 0; return 0; }
    a.cx = a.nx - 1; a.cy = a.ny - 1; a.cz = a.nz - 1;
    // tile height: ~4096 cells per tile, staged rows must fit in shared memory
    uint32_t R = 4096u / a.cx;
    if (R < 1) R = 1;
    if (R > a.cy) R = a.cy;
    const size_t fixed_bytes = 256 * 8 + 16 + 8 + 8 + 32 + kWarps * kQueue * 4 + kWarps * 8 + 256 + 256 /*slack*/;
    auto smem_for = [&](uint32_t r) {
        const size_t stride = (((size_t)(r + 1) * a.nx) + 15) & ~(size_t)15;
        return stride * 4 * 2 + stride * 2 + (size_t)r * a.cx * (a.mode == M_REGION ? 2 : 1) + 16 + fixed_bytes;
    };
    while (R > 1 && smem_for(R) > 100 * 1024) --R;
    const size_t smem = smem_for(R);
    if (smem > 227 * 1024) return fail_msg(c, "grid row too wide for one shared-memory tile (nx too large)");
    if ((size_t)R * a.cx >= (1u << 28)) return fail_msg(c, "tile too large");
    a.rows_per_tile = R;
    a.tiles_per_slice = (a.cy + R - 1) / R;
    const unsigned long long nt = (unsigned long long)a.tiles_per_slice * a.cz;
    if (nt >= 0xffffffffull) return fail_msg(c, "too many tiles");
    a.num_tiles = (uint32_t)nt;
    a.prow_stride = (uint32_t)((((size_t)(R + 1) * a.nx) + 15) & ~(size_t)15);
    {   // c / cx for c < 2^28 as (c * mul) >> shift  (round-up method: mul = ceil(2^shift / cx), shift = 28 + ceil(log2 cx))
        uint32_t l = 0;
        while ((1u << l) < a.cx) ++l;
        a.magic_cx_shift = 28 + l;
        a.magic_cx_mul = (uint32_t)(((1ull << a.magic_cx_shift) + a.cx - 1) / a.cx);
        float thr = (float)0.0005;
        if ((double)thr < 0.0005) thr = nextafterf(thr, 1.0f);
        a.snap_thr = thr;
    }
    // TMA bulk copies need 16-byte aligned global addresses and sizes
    a.use_tma = !(c->options & GCB_OPT_NO_TMA) && a.f0 && (a.nx % 4 == 0) && (((uintptr_t)a.f0 & 15) == 0);

    if (c->status_cap < a.num_tiles) {
        if (c->d_status) cudaFree(c->d_status);
        c->status_cap = (size_t)a.num_tiles + 1024;
        GCB_CHECK(c, cudaMalloc(&c->d_status, c->status_cap * 2 * sizeof(unsigned long long)));
    }
    a.status_a = c->d_status;
    a.status_v = c->d_status + c->status_cap;
    a.tile_counter = c->d_tile_counter;
    a.totals = c->d_totals;
    if (!a.count_only) {
        GCB_CHECK(c, cudaMemsetAsync(a.status_a, 0, (size_t)a.num_tiles * 8, c->stream));
        GCB_CHECK(c, cudaMemsetAsync(a.status_v, 0, (size_t)a.num_tiles * 8, c->stream));
    }
    GCB_CHECK(c, cudaMemsetAsync(c->d_tile_counter, 0, sizeof(uint32_t), c->stream));
    GCB_CHECK(c, cudaMemsetAsync(c->d_totals, 0, 2 * sizeof(unsigned long long), c->stream));

    int occ = 1;
    switch (a.mode) {
    case M_LATTICE_ONE: occ = occupancy<M_LATTICE_ONE>(smem); break;
    case M_LATTICE: occ = occupancy<M_LATTICE>(smem); break;
    case M_CSG: occ = occupancy<M_CSG>(smem); break;
    case M_TOPO: occ = occupancy<M_TOPO>(smem); break;
    case M_BAND_RAW: occ = occupancy<M_BAND_RAW>(smem); break;
    case M_REGION: occ = occupancy<M_REGION>(smem); break;
    default: return fail_msg(c, "bad mode");
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
    GCB_CHECK(c, cudaMemcpyAsync(c->h_totals, c->d_totals, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    GCB_CHECK(c, cudaStreamSynchronize(c->stream));
    *active = c->h_totals[0];
    *verts = c->h_totals[0] ? c->h_totals[1] : 0;  // early-out of Isosurface.cu:83-87
    return 0;
}

} // namespace gcb
