// Exhaustive check of the reciprocal-based division used when staging the band field (mc_extract.cu, div_by_uniform):
// for a set of divisors d, every one of the 2^32 float bit patterns n is divided both ways and compared bit for bit with
// __fdiv_rn(n, d).  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/div_check.cu -o gpurun_out/div_check
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

__device__ __forceinline__ float div_by_uniform(float n, float d, float y) {
    const float q0 = __fmul_rn(n, y);
    const float r0 = __fmaf_rn(-q0, d, n);
    const float q1 = __fmaf_rn(r0, y, q0);
    const float r1 = __fmaf_rn(-q1, d, n);
    const float q2 = __fmaf_rn(r1, y, q1);
    // tiny numerators (the exact remainder would underflow), zero (sign), tiny/huge quotients, inf/nan: IEEE division
    if (!(fabsf(n) >= 1.0e-30f && fabsf(q2) >= 1.0e-30f && fabsf(q2) <= 1.0e30f)) return __fdiv_rn(n, d);
    return q2;
}

__global__ void check(float d, unsigned long long* bad, unsigned* first_bad) {
    const float y = __frcp_rn(d);
    unsigned long long local = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < (1ull << 32); i += (unsigned long long)gridDim.x * blockDim.x) {
        const float n = __uint_as_float((unsigned)i);
        const float a = div_by_uniform(n, d, y), b = __fdiv_rn(n, d);
        const bool same = (__float_as_uint(a) == __float_as_uint(b)) || (a != a && b != b);
        if (!same) { ++local; atomicMin(first_bad, (unsigned)i); }
    }
    if (local) atomicAdd(bad, local);
}

int main(int argc, char** argv) {
    const int nrand = argc > 1 ? atoi(argv[1]) : 64;
    unsigned long long* bad; unsigned* fb;
    cudaMalloc(&bad, 8); cudaMalloc(&fb, 4);
    unsigned specials[] = {0x3f800000u, 0x3fffffffu, 0x3f800001u, 0x407fffffu, 0x3effffffu, 0x41200000u, 0x3dcccccdu, 0x42c80000u, 0x3a83126fu,
                           0x4b000000u, 0x33800000u, 0x5d000000u, 0x21000000u, 0xbf800000u, 0xc0490fdbu, 0x7f7fffffu, 0x00800000u, 0x00000001u};
    const int nspec = sizeof(specials) / sizeof(specials[0]);
    srand(12345);
    unsigned long long total_bad = 0;
    for (int t = 0; t < nspec + nrand; ++t) {
        unsigned bits;
        if (t < nspec) bits = specials[t];
        else {
            // random mantissa, exponent in the range the normaliser sees (2^-40 .. 2^40), either sign
            const unsigned man = ((unsigned)rand() << 8 ^ (unsigned)rand()) & 0x7fffffu;
            const unsigned ex = 127 - 40 + (unsigned)(rand() % 81);
            bits = ((rand() & 7) == 0 ? 0x80000000u : 0u) | (ex << 23) | man;
        }
        float d; memcpy(&d, &bits, 4);
        cudaMemset(bad, 0, 8); cudaMemset(fb, 0xff, 4);
        check<<<148 * 16, 256>>>(d, bad, fb);
        unsigned long long hb; unsigned hf;
        cudaMemcpy(&hb, bad, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&hf, fb, 4, cudaMemcpyDeviceToHost);
        total_bad += hb;
        if (hb || t < nspec) printf("d = %.9g (0x%08x): %llu mismatches%s", d, bits, hb, hb ? "" : "\n");
        if (hb) printf(", first n = 0x%08x\n", hf);
    }
    printf("divisors tested: %d, total mismatches: %llu\n", nspec + nrand, total_bad);
    return total_bad ? 1 : 0;
}
