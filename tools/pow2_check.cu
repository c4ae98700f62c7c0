// pow2_check.cu -- is libdevice powf(x, 2) the same float as x * x for EVERY x?  (exhaustive over all 2^32 bit patterns)
// The reference's primitive kernels write `powf(v, 2)` (Modelling.cu:266-745); nvcc does not fold it into a product, so every point
// pays three ~65-instruction pow evaluations.  If the two agree bit for bit (NaNs compared as NaNs) the product is a drop-in.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/pow2_check tools/pow2_check.cu && tools/bin/pow2_check
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void check(unsigned long long* bad, uint32_t* first, unsigned long long* bad_sq) {
    const uint64_t n = 1ull << 32;
    unsigned long long local = 0, local2 = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((uint32_t)i);
        const float a = powf(x, 2);          // as the reference spells it
        const float b = __fmul_rn(x, x);
        const bool same = (a != a && b != b) || __float_as_uint(a) == __float_as_uint(b);
        if (!same) {
            ++local;
            if (fabsf(b) >= 1.17549435e-38f || b == 0.0f || b != b) atomicAdd(bad + 2, 1ull);  // mismatch although x*x is normal, zero, inf or nan
            const unsigned long long k = atomicAdd(bad + 1, 1ull);
            if (k < 16) first[k] = (uint32_t)i;
        }
        // sqrtf(powf(x,2) + ...) style users only see the value, nothing else to check; also compare pow(x, 2) in double (cone frustum)
        const float c = (float)pow((double)x, 2.0);  // not used by the replacement, informational
        if (!((c != c && b != b) || __float_as_uint(c) == __float_as_uint(b))) ++local2;
    }
    atomicAdd(bad, local);
    atomicAdd(bad_sq, local2);
}

int main() {
    unsigned long long *d_bad, *d_bad2, h[3] = {0, 0, 0}, h2 = 0;
    uint32_t *d_first, hf[16] = {0};
    cudaMalloc(&d_bad, 24); cudaMalloc(&d_bad2, 8); cudaMalloc(&d_first, 64);
    cudaMemset(d_bad, 0, 24); cudaMemset(d_bad2, 0, 8); cudaMemset(d_first, 0, 64);
    check<<<148 * 16, 256>>>(d_bad, d_first, d_bad2);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d_bad, 24, cudaMemcpyDeviceToHost); cudaMemcpy(&h2, d_bad2, 8, cudaMemcpyDeviceToHost); cudaMemcpy(hf, d_first, 64, cudaMemcpyDeviceToHost);
    printf("{\"status\": \"%s\", \"inputs\": 4294967296, \"powf_x_2_vs_fmul_mismatches\": %llu, \"of_which_product_not_denormal\": %llu, \"float_of_double_pow_vs_fmul_mismatches\": %llu, \"first\": [", cudaGetErrorString(e), h[0], h[2], h2);
    for (int i = 0; i < 16 && i < (int)h[1]; ++i) printf("%s\"0x%08x\"", i ? ", " : "", hf[i]);
    printf("]}\n");
    return 0;
}
