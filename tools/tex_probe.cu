// tex_probe.cu -- which software formula reproduces tex3D<float> linear filtering bit for bit?
// Diagnostic only (run once on the GPU box; results recorded in DESIGN.md).  Samples a random float
// 3-D texture exactly as the reference does (Gratings.cu:676: unnormalised coords idx*d + 0.5) for
// upsampling ratios 2, 4, 8 and compares candidate combination orders against the texture unit.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

#define NC 7
struct Axis { int i0, i1; float a; };
__device__ Axis tex_axis(float coord, int n) {
    float xb = coord - 0.5f, fl = floorf(xb);
    float a = rintf((xb - fl) * 256.0f) * (1.0f / 256.0f);
    int i = (int)fl;
    if (a >= 1.0f) { a = 0.f; i += 1; }
    Axis r; r.i0 = min(max(i, 0), n - 1); r.i1 = min(max(i + 1, 0), n - 1); r.a = a; return r;
}
__device__ float lerp_fma(float t0, float t1, float a) { return __fmaf_rn(a, __fsub_rn(t1, t0), t0); }
__device__ float lerp_w(float t0, float t1, float a) { return __fmaf_rn(a, t1, __fmul_rn(__fsub_rn(1.f, a), t0)); }
__device__ float lerp_nofma(float t0, float t1, float a) { return __fadd_rn(t0, __fmul_rn(a, __fsub_rn(t1, t0))); }

template <int C> __device__ float combine(const float t[2][2][2], float a, float b, float c) {
    if (C == 0 || C == 1 || C == 5) {
        auto L = [](float p, float q, float w) { return C == 0 ? lerp_fma(p, q, w) : C == 1 ? lerp_w(p, q, w) : lerp_nofma(p, q, w); };
        float x00 = L(t[0][0][0], t[0][0][1], a), x01 = L(t[0][1][0], t[0][1][1], a), x10 = L(t[1][0][0], t[1][0][1], a), x11 = L(t[1][1][0], t[1][1][1], a);
        float y0 = L(x00, x01, b), y1 = L(x10, x11, b);
        return L(y0, y1, c);
    } else if (C == 2) {
        double r = 0;
        for (int k = 0; k < 2; ++k) for (int j = 0; j < 2; ++j) for (int i = 0; i < 2; ++i)
            r += (double)(i ? a : 1 - a) * (double)(j ? b : 1 - b) * (double)(k ? c : 1 - c) * (double)t[k][j][i];
        return (float)r;
    } else if (C == 3) {
        float r = 0;
        for (int k = 0; k < 2; ++k) for (int j = 0; j < 2; ++j) for (int i = 0; i < 2; ++i)
            r = __fmaf_rn(__fmul_rn(__fmul_rn(i ? a : 1 - a, j ? b : 1 - b), k ? c : 1 - c), t[k][j][i], r);
        return r;
    } else if (C == 4) {  // z first, then y, then x
        float z00 = lerp_fma(t[0][0][0], t[1][0][0], c), z01 = lerp_fma(t[0][0][1], t[1][0][1], c), z10 = lerp_fma(t[0][1][0], t[1][1][0], c), z11 = lerp_fma(t[0][1][1], t[1][1][1], c);
        float y0 = lerp_fma(z00, z10, b), y1 = lerp_fma(z01, z11, b);
        return lerp_fma(y0, y1, a);
    } else {  // C == 6: double lerps x,y,z with single final rounding
        double x00 = t[0][0][0] + (double)a * ((double)t[0][0][1] - t[0][0][0]), x01 = t[0][1][0] + (double)a * ((double)t[0][1][1] - t[0][1][0]);
        double x10 = t[1][0][0] + (double)a * ((double)t[1][0][1] - t[1][0][0]), x11 = t[1][1][0] + (double)a * ((double)t[1][1][1] - t[1][1][0]);
        double y0 = x00 + (double)b * (x01 - x00), y1 = x10 + (double)b * (x11 - x10);
        return (float)(y0 + (double)c * (y1 - y0));
    }
}

__global__ void probe(cudaTextureObject_t tex, const float* g, int cx, int cy, int cz, int nx, int ny, int nz, float d, unsigned long long* bad, int* maxulp) {
    size_t n = (size_t)nx * ny * nz;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int tx = i % nx, ty = (i / nx) % ny, tz = i / ((size_t)nx * ny);
        float x = tx * d, y = ty * d, z = tz * d;
        float X = (float)(x + 0.5), Y = (float)(y + 0.5), Z = (float)(z + 0.5);
        float hw = tex3D<float>(tex, X, Y, Z);
        Axis ax = tex_axis(X, cx), ay = tex_axis(Y, cy), az = tex_axis(Z, cz);
        float t[2][2][2];
        int zi[2] = {az.i0, az.i1}, yi[2] = {ay.i0, ay.i1}, xi[2] = {ax.i0, ax.i1};
        for (int k = 0; k < 2; ++k) for (int j = 0; j < 2; ++j) for (int ii = 0; ii < 2; ++ii) t[k][j][ii] = g[((size_t)zi[k] * cy + yi[j]) * cx + xi[ii]];
        float c[NC] = {combine<0>(t, ax.a, ay.a, az.a), combine<1>(t, ax.a, ay.a, az.a), combine<2>(t, ax.a, ay.a, az.a), combine<3>(t, ax.a, ay.a, az.a),
                       combine<4>(t, ax.a, ay.a, az.a), combine<5>(t, ax.a, ay.a, az.a), combine<6>(t, ax.a, ay.a, az.a)};
        for (int q = 0; q < NC; ++q) {
            int u = abs(__float_as_int(c[q]) - __float_as_int(hw));
            if (u) { atomicAdd(bad + q, 1ull); atomicMax(maxulp + q, u); }
        }
    }
}

// dump raw samples {8 taps, ax, ay, az, hw} for offline analysis of the filter arithmetic
__global__ void dump(cudaTextureObject_t tex, const float* g, int cx, int cy, int cz, int nx, int ny, int nz, float d, float* out, int nsamp) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsamp) return;
    unsigned long long h = (unsigned long long)(s + 1) * 0x9E3779B97F4A7C15ull;
    int tx = (h >> 8) % nx, ty = (h >> 24) % ny, tz = (h >> 40) % nz;
    float x = tx * d, y = ty * d, z = tz * d;
    float X = (float)(x + 0.5), Y = (float)(y + 0.5), Z = (float)(z + 0.5);
    Axis ax = tex_axis(X, cx), ay = tex_axis(Y, cy), az = tex_axis(Z, cz);
    int zi[2] = {az.i0, az.i1}, yi[2] = {ay.i0, ay.i1}, xi[2] = {ax.i0, ax.i1};
    float* o = out + (size_t)s * 12;
    int q = 0;
    for (int k = 0; k < 2; ++k) for (int j = 0; j < 2; ++j) for (int i = 0; i < 2; ++i) o[q++] = g[((size_t)zi[k] * cy + yi[j]) * cx + xi[i]];
    o[8] = ax.a; o[9] = ay.a; o[10] = az.a; o[11] = tex3D<float>(tex, X, Y, Z);
}

int main(int argc, char** argv) {
    const int cx = 40, cy = 36, cz = 32;
    std::vector<float> h((size_t)cx * cy * cz);
    srand(1);
    const bool wild = argc > 2;  // values spread over 2^24 in magnitude (exercise alignment / truncation of the filter)
    for (auto& v : h) { v = ((rand() % 200001) - 100000) * 0.00123f; if (wild) v = ldexpf(v, -(rand() % 24)); }
    float* dg; cudaMalloc(&dg, h.size() * 4); cudaMemcpy(dg, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaArray_t arr; cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
    cudaExtent ext = make_cudaExtent(cx, cy, cz);
    cudaMalloc3DArray(&arr, &desc, ext);
    cudaMemcpy3DParms p = {0};
    p.srcPtr = make_cudaPitchedPtr(dg, cx * 4, cx, cy); p.dstArray = arr; p.extent = ext; p.kind = cudaMemcpyDeviceToDevice;
    cudaMemcpy3D(&p);
    cudaResourceDesc rd; memset(&rd, 0, sizeof rd); rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
    cudaTextureDesc td; memset(&td, 0, sizeof td); td.normalizedCoords = false; td.filterMode = cudaFilterModeLinear; td.addressMode[0] = cudaAddressModeWrap; td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex; cudaCreateTextureObject(&tex, &rd, &td, NULL);
    unsigned long long* bad; int* mu; cudaMalloc(&bad, NC * 8); cudaMalloc(&mu, NC * 4);
    const char* names[NC] = {"sep xyz fma(a,t1-t0,t0)", "sep xyz fma(a,t1,(1-a)t0)", "8-tap double sum", "8-tap float fma sum", "sep zyx fma", "sep xyz mul+add", "sep xyz double"};
    for (int ratio : {2, 4, 8, 3}) {
        float d = 1.0f / ratio;
        int nx = cx * ratio, ny = cy * ratio, nz = cz * ratio;
        cudaMemset(bad, 0, NC * 8); cudaMemset(mu, 0, NC * 4);
        probe<<<1024, 256>>>(tex, dg, cx, cy, cz, nx, ny, nz, d, bad, mu);
        unsigned long long hb[NC]; int hm[NC];
        cudaMemcpy(hb, bad, NC * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hm, mu, NC * 4, cudaMemcpyDeviceToHost);
        printf("ratio %d (%d x %d x %d points): %s\n", ratio, nx, ny, nz, cudaGetErrorString(cudaGetLastError()));
        for (int q = 0; q < NC; ++q) printf("   cand %d %-28s mismatches %12llu  max ulp %d\n", q, names[q], hb[q], hm[q]);
    }
    if (argc > 1) {
        const int nsamp = 200000;
        float* dout; cudaMalloc(&dout, (size_t)nsamp * 12 * 4);
        std::vector<float> ho((size_t)nsamp * 12);
        for (int ratio : {2, 4, 8}) {
            dump<<<(nsamp + 255) / 256, 256>>>(tex, dg, cx, cy, cz, cx * ratio, cy * ratio, cz * ratio, 1.0f / ratio, dout, nsamp);
            cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
            char name[512]; snprintf(name, sizeof name, "%s/tex_samples%s_r%d.bin", argv[1], wild ? "_wild" : "", ratio);
            FILE* f = fopen(name, "wb"); fwrite(ho.data(), 4, ho.size(), f); fclose(f);
        }
    }
    return 0;
}
