// obj_gpu.cu -- .obj export on the GPU with the reference's exact file bytes.
//
// Reference: File_output::file_write_obj, src/File_output.cu:5-81 -- a single host thread that quantises every vertex,
// welds through std::map<std::vector<float>,int> (one heap allocation and an O(log V) lexicographic compare per vertex),
// filters faces through a second map and streams text.  Observable behaviour kept here (same rules as obj_writer.cpp, the
// host restatement this file is tested against byte for byte):
//   * v = float(int(p * 1000) * 0.001) per component (:26-28);
//   * vertices are welded on the quantised triple, the FIRST occurrence defines the 1-based id (:30-48);
//   * a face is dropped when two of its ids coincide or when the same ordered id triple was already written (:63-75);
//     winding is flipped on output, " f  a c b" (:72);
//   * numbers are printed with iostream defaults (== "%g").
// Data-parallel formulation:
//   1. quantise                          -> 3 x u32 key per vertex
//   2. weld   : open-addressing hash whose slots hold a VERTEX INDEX; equal keys meet in one slot and atomicMin keeps the
//               smallest index = first occurrence (keys are compared through the representative's key, so a slot needs no
//               96-bit CAS)
//   3. ids    : exclusive scan of the "is first occurrence" flags; id(i) = scan[first(i)] + 1
//   4. faces  : same hash on the ordered id triple over non-degenerate triangles, keep the first occurrence
//   5. text   : per-line byte counts -> exclusive scan -> every kept vertex / face formats its line at its offset
//   6. D2H of the text in chunks, fwrite.
// The scans are cub::DeviceScan (CUDA toolkit), everything else is in this file.
#include "common.cuh"

#include <cub/device/device_scan.cuh>

#include <cstdio>
#include <vector>

namespace gcb {

namespace {

constexpr uint32_t kEmpty = 0xffffffffu;

__device__ __forceinline__ uint64_t mix3(uint32_t a, uint32_t b, uint32_t c) {
    uint64_t h = a * 0x9E3779B97F4A7C15ull;
    h ^= (h >> 29);
    h += b * 0xBF58476D1CE4E5B9ull;
    h ^= (h >> 32);
    h += c * 0x94D049BB133111EBull;
    h ^= (h >> 31);
    return h * 0xD6E8FEB86659FD93ull;
}

// File_output.cu:26-28: `float vx = int(p.x * 1000) * 0.001;` -- float product, truncation, double product, narrowing
__device__ __forceinline__ int quant_int(float p) { return (int)__fmul_rn(p, 1000.0f); }
__device__ __forceinline__ float quant(float p) { return __double2float_rn(__dmul_rn((double)quant_int(p), 0.001)); }
__device__ __forceinline__ uint32_t key_bits(float v) { return __float_as_uint(v == 0.0f ? 0.0f : v); }  // the map treats -0 == +0

struct Key3 { uint32_t a, b, c; };
__device__ __forceinline__ bool same(const Key3& x, const Key3& y) { return x.a == y.a && x.b == y.b && x.c == y.c; }

__global__ void __launch_bounds__(256) quantise_kernel(const float4* __restrict__ pos, uint32_t n, Key3* __restrict__ keys) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = pos[i];
        keys[i] = Key3{key_bits(quant(p.x)), key_bits(quant(p.y)), key_bits(quant(p.z))};
    }
}

// insert item i (key keys[i]) into a table of item indices; on return the slot of i's key holds min(index) of that key
__device__ __forceinline__ void hash_insert(uint32_t* table, uint64_t mask, const Key3* keys, uint32_t i) {
    const Key3 k = keys[i];
    uint64_t h = (mix3(k.a, k.b, k.c) >> 7) & mask;
    for (;;) {
        uint32_t cur = table[h];
        if (cur == kEmpty) {
            cur = atomicCAS(&table[h], kEmpty, i);
            if (cur == kEmpty) return;
        }
        if (same(keys[cur], k)) {  // the representative may change under us, its key can not
            atomicMin(&table[h], i);
            return;
        }
        h = (h + 1) & mask;
    }
}
__device__ __forceinline__ uint32_t hash_find(const uint32_t* table, uint64_t mask, const Key3* keys, uint32_t i) {
    const Key3 k = keys[i];
    uint64_t h = (mix3(k.a, k.b, k.c) >> 7) & mask;
    for (;;) {
        const uint32_t cur = table[h];
        if (same(keys[cur], k)) return cur;  // every key was inserted: no empty slot before it
        h = (h + 1) & mask;
    }
}

__global__ void __launch_bounds__(256) insert_kernel(uint32_t* table, uint64_t mask, const Key3* __restrict__ keys, uint32_t n, const unsigned char* __restrict__ valid) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (!valid || valid[i]) hash_insert(table, mask, keys, i);
}
// first[i] = index of the first occurrence of i's key; flag[i] = (first[i] == i)
__global__ void __launch_bounds__(256) first_kernel(const uint32_t* __restrict__ table, uint64_t mask, const Key3* __restrict__ keys, uint32_t n,
                                                    const unsigned char* __restrict__ valid, uint32_t* __restrict__ first, uint32_t* __restrict__ flag) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t f = kEmpty;
        if (!valid || valid[i]) f = hash_find(table, mask, keys, i);
        if (first) first[i] = f;
        flag[i] = f == i ? 1u : 0u;
    }
}
// ids (1-based) of the three corners of every triangle; degenerate triangles are marked invalid
__global__ void __launch_bounds__(256) face_keys_kernel(const uint32_t* __restrict__ first, const uint32_t* __restrict__ vscan, uint32_t ntri, Key3* __restrict__ fkeys,
                                                        unsigned char* __restrict__ valid) {
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < ntri; t += gridDim.x * blockDim.x) {
        const uint32_t a = vscan[first[3 * t]] + 1u, b = vscan[first[3 * t + 1]] + 1u, c = vscan[first[3 * t + 2]] + 1u;
        fkeys[t] = Key3{a, b, c};
        valid[t] = (a != b && a != c && b != c) ? 1 : 0;
    }
}

// ---- "%g" (6 significant digits, iostream default) for the floats this writer prints.  A nonzero quantised coordinate is
// at least 0.001 in magnitude; scaling |v| (24 significant bits) by 10^n, n <= 8, is exact in double, so the digits below are
// the correctly rounded ones (round-half-even on the exact value, as glibc prints) for |v| < 10^6.  Larger magnitudes use the
// same code with one inexact division (a tie there would need a coordinate beyond 10^6, which int(p * 1000) cannot carry).
__device__ int fmt_g(float vf, char* out) {  // returns the length; out == nullptr: count only
    int n = 0;
    auto put = [&](char ch) { if (out) out[n] = ch; ++n; };
    double a = (double)vf;
    if (a == 0.0) { if (signbit(vf)) put('-'); put('0'); return n; }
    if (a < 0) { put('-'); a = -a; }
    int e10 = 0;
    double p10 = 1.0;
    while (a >= p10 * 10.0) { p10 *= 10.0; ++e10; }
    while (a < p10) { p10 /= 10.0; --e10; }
    // D = a / 10^(e10-5) rounded to an integer in [10^5, 10^6]
    double scaled;
    if (e10 <= 5) { double m = 1.0; for (int k = e10; k < 5; ++k) m *= 10.0; scaled = a * m; }
    else { double m = 1.0; for (int k = 5; k < e10; ++k) m *= 10.0; scaled = a / m; }
    long long D = (long long)rint(scaled);
    if (D >= 1000000) { D = 100000; ++e10; }
    char dig[6];
    for (int k = 5; k >= 0; --k) { dig[k] = (char)('0' + (int)(D % 10)); D /= 10; }
    int nd = 6;
    while (nd > 1 && dig[nd - 1] == '0') --nd;  // significant digits after stripping trailing zeros
    if (e10 < -4 || e10 >= 6) {
        put(dig[0]);
        if (nd > 1) { put('.'); for (int k = 1; k < nd; ++k) put(dig[k]); }
        put('e'); put(e10 < 0 ? '-' : '+');
        const int ae = e10 < 0 ? -e10 : e10;
        put((char)('0' + ae / 10)); put((char)('0' + ae % 10));
    } else if (e10 >= 0) {
        for (int k = 0; k <= e10; ++k) put(k < nd ? dig[k] : '0');
        if (nd > e10 + 1) { put('.'); for (int k = e10 + 1; k < nd; ++k) put(dig[k]); }
    } else {
        put('0'); put('.');
        for (int k = -1; k > e10; --k) put('0');
        for (int k = 0; k < nd; ++k) put(dig[k]);
    }
    return n;
}
__device__ int fmt_u(uint32_t v, char* out) {
    char tmp[10];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10u); v /= 10u; } while (v);
    if (out) for (int k = 0; k < n; ++k) out[k] = tmp[n - 1 - k];
    return n;
}
__device__ int vertex_line(float4 p, char* out) {  // "v x y z\n"
    int n = 0;
    if (out) { out[0] = 'v'; out[1] = ' '; }
    n = 2;
    n += fmt_g(quant(p.x), out ? out + n : nullptr);
    if (out) out[n] = ' ';
    ++n;
    n += fmt_g(quant(p.y), out ? out + n : nullptr);
    if (out) out[n] = ' ';
    ++n;
    n += fmt_g(quant(p.z), out ? out + n : nullptr);
    if (out) out[n] = '\n';
    return n + 1;
}
__device__ int face_line(Key3 f, char* out) {  // " f  a c b\n"
    int n = 4;
    if (out) { out[0] = ' '; out[1] = 'f'; out[2] = ' '; out[3] = ' '; }
    n += fmt_u(f.a, out ? out + n : nullptr);
    if (out) out[n] = ' ';
    ++n;
    n += fmt_u(f.c, out ? out + n : nullptr);
    if (out) out[n] = ' ';
    ++n;
    n += fmt_u(f.b, out ? out + n : nullptr);
    if (out) out[n] = '\n';
    return n + 1;
}
__global__ void __launch_bounds__(256) vertex_len_kernel(const float4* __restrict__ pos, const uint32_t* __restrict__ flag, uint32_t n, uint32_t* __restrict__ len) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) len[i] = flag[i] ? (uint32_t)vertex_line(pos[i], nullptr) : 0u;
}
__global__ void __launch_bounds__(256) vertex_text_kernel(const float4* __restrict__ pos, const uint32_t* __restrict__ flag, const unsigned long long* __restrict__ off,
                                                          uint32_t n, char* __restrict__ text) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (flag[i]) vertex_line(pos[i], text + off[i]);
}
__global__ void __launch_bounds__(256) face_len_kernel(const Key3* __restrict__ fkeys, const uint32_t* __restrict__ keep, uint32_t n, uint32_t* __restrict__ len) {
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) len[t] = keep[t] ? (uint32_t)face_line(fkeys[t], nullptr) : 0u;
}
__global__ void __launch_bounds__(256) face_text_kernel(const Key3* __restrict__ fkeys, const uint32_t* __restrict__ keep, const unsigned long long* __restrict__ off,
                                                        uint32_t n, char* __restrict__ text) {
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x)
        if (keep[t]) face_line(fkeys[t], text + off[t]);
}
__global__ void widen_kernel(const uint32_t* __restrict__ in, uint32_t n, unsigned long long* __restrict__ out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = in[i];
}

struct DevBuf {  // RAII for the temporaries of one export
    std::vector<void*> ptrs;
    ~DevBuf() { for (void* p : ptrs) cudaFree(p); }
    template <class T>
    cudaError_t get(T** p, size_t count) {
        cudaError_t e = cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
};

template <class In, class Out>
cudaError_t exclusive_sum(DevBuf& mem, const In* in, Out* out, uint32_t n, cudaStream_t st) {
    size_t bytes = 0;
    cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, st);
    if (e != cudaSuccess) return e;
    char* tmp;
    if ((e = mem.get(&tmp, bytes)) != cudaSuccess) return e;
    return cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, (int)n, st);
}

} // namespace

#define OBJ_CHECK(call)                                                                   \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) { if (f) fclose(f); return fail(c, "file_write_obj (device): " #call, e_); } \
    } while (0)

int write_obj_device(Ctx* c, const float4* pos, unsigned int total_verts, const char* filename) {
    FILE* f = fopen(filename, "wb");
    if (!f) return fail_msg(c, std::string("file_write_obj: cannot write ") + filename);
    static const char header[] = "##Sample latttice new Obj \no Solid \n";
    bool ok = fwrite(header, 1, sizeof(header) - 1, f) == sizeof(header) - 1;
    const uint32_t nv = total_verts, nt = total_verts / 3;
    cudaStream_t st = c->stream;
    const unsigned blocks = (unsigned)c->num_sms * 8;
    DevBuf mem;
    unsigned long long v_bytes = 0, f_bytes = 0;
    char* text = nullptr;
    unsigned long long *voff = nullptr, *foff = nullptr;
    if (nv) {
        // ---- weld
        Key3* keys; uint32_t *table, *first, *flag, *vscan, *len;
        uint64_t cap = 16;
        while (cap < 2ull * nv) cap <<= 1;
        OBJ_CHECK(mem.get(&keys, nv));
        OBJ_CHECK(mem.get(&table, cap));
        OBJ_CHECK(mem.get(&first, nv));
        OBJ_CHECK(mem.get(&flag, nv));
        OBJ_CHECK(mem.get(&vscan, nv));
        OBJ_CHECK(mem.get(&len, std::max(nv, nt)));
        OBJ_CHECK(mem.get(&voff, nv));
        quantise_kernel<<<blocks, 256, 0, st>>>(pos, nv, keys);
        OBJ_CHECK(cudaMemsetAsync(table, 0xff, cap * sizeof(uint32_t), st));
        insert_kernel<<<blocks, 256, 0, st>>>(table, cap - 1, keys, nv, nullptr);
        first_kernel<<<blocks, 256, 0, st>>>(table, cap - 1, keys, nv, nullptr, first, flag);
        OBJ_CHECK(exclusive_sum(mem, flag, vscan, nv, st));
        // ---- vertex text offsets
        vertex_len_kernel<<<blocks, 256, 0, st>>>(pos, flag, nv, len);
        widen_kernel<<<blocks, 256, 0, st>>>(len, nv, voff);
        OBJ_CHECK(exclusive_sum(mem, voff, voff, nv, st));
        unsigned long long last_off = 0; uint32_t last_len = 0;
        OBJ_CHECK(cudaMemcpyAsync(&last_off, voff + nv - 1, 8, cudaMemcpyDeviceToHost, st));
        OBJ_CHECK(cudaMemcpyAsync(&last_len, len + nv - 1, 4, cudaMemcpyDeviceToHost, st));
        OBJ_CHECK(cudaStreamSynchronize(st));
        v_bytes = last_off + last_len;
        // ---- faces
        Key3* fkeys = nullptr; unsigned char* valid = nullptr; uint32_t *ftable = nullptr, *keep = nullptr;
        if (nt) {
            uint64_t fcap = 16;
            while (fcap < 2ull * nt) fcap <<= 1;
            OBJ_CHECK(mem.get(&fkeys, nt));
            OBJ_CHECK(mem.get(&valid, nt));
            OBJ_CHECK(mem.get(&ftable, fcap));
            OBJ_CHECK(mem.get(&keep, nt));
            OBJ_CHECK(mem.get(&foff, nt));
            face_keys_kernel<<<blocks, 256, 0, st>>>(first, vscan, nt, fkeys, valid);
            OBJ_CHECK(cudaMemsetAsync(ftable, 0xff, fcap * sizeof(uint32_t), st));
            insert_kernel<<<blocks, 256, 0, st>>>(ftable, fcap - 1, fkeys, nt, valid);
            first_kernel<<<blocks, 256, 0, st>>>(ftable, fcap - 1, fkeys, nt, valid, nullptr, keep);
            face_len_kernel<<<blocks, 256, 0, st>>>(fkeys, keep, nt, len);
            widen_kernel<<<blocks, 256, 0, st>>>(len, nt, foff);
            OBJ_CHECK(exclusive_sum(mem, foff, foff, nt, st));
            OBJ_CHECK(cudaMemcpyAsync(&last_off, foff + nt - 1, 8, cudaMemcpyDeviceToHost, st));
            OBJ_CHECK(cudaMemcpyAsync(&last_len, len + nt - 1, 4, cudaMemcpyDeviceToHost, st));
            OBJ_CHECK(cudaStreamSynchronize(st));
            f_bytes = last_off + last_len;
        }
        // ---- text: vertices, "\n\n", faces
        const unsigned long long total = v_bytes + 2 + f_bytes;
        OBJ_CHECK(mem.get(&text, total));
        vertex_text_kernel<<<blocks, 256, 0, st>>>(pos, flag, voff, nv, text);
        OBJ_CHECK(cudaMemsetAsync(text + v_bytes, '\n', 2, st));
        if (nt) face_text_kernel<<<blocks, 256, 0, st>>>(fkeys, keep, foff, nt, text + v_bytes + 2);
        c->launches += nt ? 13 : 7;
        OBJ_CHECK(cudaGetLastError());
        // ---- D2H in chunks through a pinned bounce buffer
        const size_t chunk = 64u << 20;
        char* bounce;
        OBJ_CHECK(cudaMallocHost(&bounce, chunk));
        for (unsigned long long o = 0; o < total && ok; o += chunk) {
            const size_t nb = (size_t)std::min<unsigned long long>(chunk, total - o);
            cudaError_t e = cudaMemcpyAsync(bounce, text + o, nb, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { cudaFreeHost(bounce); fclose(f); return fail(c, "file_write_obj (device): text D2H", e); }
            ok = fwrite(bounce, 1, nb, f) == nb;
        }
        cudaFreeHost(bounce);
    } else {
        ok = ok && fwrite("\n\n", 1, 2, f) == 2;
    }
    if (fclose(f) != 0 || !ok) return fail_msg(c, std::string("file_write_obj: cannot write ") + filename);
    return 0;
}

} // namespace gcb
