// phase_solve.cu -- the phase solve of the spatially varying lattice.
//
// Reference (Multitopo::spatial_lattice_run, main.cu:3949-3962): for each of the 62 harmonics (i, j, k)
//     finding_phi   (Gratings.cu:100-417, :1015-1025)  right-hand side  b = D^T K  of the least-squares problem grad(phi) = K,
//                                                      K the locally rotated / rescaled grating vector
//     GPUCG_lattice (Gratings.cu:875-974)              plain CG on  (D^T D) phi = b  with the 7-band stencil of
//                                                      GPUMatvec_lattice_kernel (:420-597), <= 500 iterations, |r| <= 0.01
// driven from the host: per iteration 4 kernels + 2 two-level reductions, 3 blocking 4-byte D2H copies, a cudaMalloc/cudaFree
// set per solve, cudaDeviceSynchronize after every launch.
//
// Here the arithmetic is the reference's, bit for bit -- same expression contraction (read from the reference SASS, confirmed
// on the GPU against the reference kernels), same reduction trees (1024-element shared-memory tree per block, then the strided
// serial sum + tree of Reduction_lattice) -- but the CG scalars live on the device and ALL harmonics advance together:
//   * one launch per CG stage covers every harmonic (blockIdx.y = harmonic); a harmonic that has converged or hit the
//     iteration limit is skipped by its blocks, so each harmonic performs exactly the reference's iterations;
//   * matvec + <d, q>, the two updates + <r, r> and the direction update are three kernels per iteration for all harmonics;
//   * the host only polls a "harmonics still running" counter every few iterations.
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

namespace gcb {

namespace {

// ---------------------------------------------------------------- right-hand side (finding_phi_kernel)
struct PhiArgs {
    int nx, ny, nz;
    float dx, dy, dz;
    int latticetype;   // 'n', 'r', 'b', 's'
    int uniform_type;  // 0 constant period, 1 per-axis periods, 2 period field
    float const_period, x_period, y_period, z_period, lcon, lcon_1;
    int sinewave_zaxis;
};

// one-sided / central difference weights of D^T at index v of n (Gratings.cu:133-253): (a1, a2) and the two sample indices
__device__ __forceinline__ void dt_weights(int v, int n, float& a1, float& a2, int& v1, int& v2) {
    if (v == 0) { a1 = -1; a2 = -0.5; v1 = v; v2 = v + 1; }
    else if (v == 1) { a1 = 1; a2 = -0.5; v1 = v - 1; v2 = v + 1; }
    else if (v == n - 2) { a1 = 0.5; a2 = -1.0; v1 = v - 1; v2 = v + 1; }
    else if (v == n - 1) { a1 = 0.5; a2 = 1; v1 = v - 1; v2 = v; }
    else { a1 = 0.5; a2 = -0.5; v1 = v - 1; v2 = v + 1; }
}

// `((2*M_PI)/per) * inner` : double division and product, narrowed to float (Gratings.cu:381-388)
__device__ __forceinline__ float k_scaled(float per, float inner) { return (float)((6.283185307179586 / (double)per) * (double)inner); }
__device__ __forceinline__ float k_scaled_int(float per, int k) { return (float)((6.283185307179586 / (double)per) * (double)k); }

// a*u + b*v in the contraction the reference build carries (checked bit for bit against its kernels, all lattice types, tools/
// phase_solve_probe.py): the first product is rounded, the second is fused
__device__ __forceinline__ float pair_sum(float a, float u, float b, float v) { return __fmaf_rn(b, v, __fmul_rn(a, u)); }

__global__ void __launch_bounds__(256) finding_phi_kernel(float* __restrict__ phi_all, const float* __restrict__ period, const int3* __restrict__ ijk, int nharm, PhiArgs A) {
    const int n = A.nx * A.ny * A.nz;
    const int h = blockIdx.y;
    const int3 f = ijk[h];
    const float fi = (float)f.x, fj = (float)f.y;
    float* phi = phi_all + (size_t)h * n;
    for (int tx = blockIdx.x * blockDim.x + threadIdx.x; tx < n; tx += gridDim.x * blockDim.x) {
        const int x = tx % A.nx, y = (tx % (A.nx * A.ny)) / A.nx, z = tx / (A.nx * A.ny);
        float a1, a2, b1, b2, c1, c2;
        int x1, x2, y1, y2, z1, z2;
        dt_weights(x, A.nx, a1, a2, x1, x2);
        dt_weights(y, A.ny, b1, b2, y1, y2);
        dt_weights(z, A.nz, c1, c2, z1, z2);
        float theta1 = 0.f, theta2 = 0.f, theta3 = 0.f, theta4 = 0.f, theta5 = 0.f, theta6 = 0.f, theta7 = 0.f, theta8 = 0.f;
        float con = 1.0f, angl = 0.0f;
        if (A.latticetype == 'r' || A.latticetype == 'b') {
            float mean_x = 0.f, mean_y = 0.f;
            if (A.latticetype == 'r') { mean_x = ((A.nx + 1) / 2.0f); mean_y = ((A.ny + 1) / 2.0f); con = 8.0f; }
            const float yy = ((y + 1) - mean_y) * A.dx;
            theta1 = atan2f(yy, ((x1 + 1) - mean_x) * A.dx);
            theta2 = atan2f(yy, ((x2 + 1) - mean_x) * A.dx);
            const float xx = ((x + 1) - mean_x) * A.dx;
            theta3 = atan2f(((y1 + 1) - mean_y) * A.dx, xx);
            theta4 = atan2f(((y2 + 1) - mean_y) * A.dx, xx);
        } else if (A.latticetype == 's') {
            const float lcon = A.lcon, lcon_1 = A.lcon_1;
            theta1 = lcon * sinf(6.28 * lcon_1 * (((x1 + 1) - 0.f) * A.dx));
            theta5 = lcon * sinf(6.28 * lcon_1 * (((z1 + 1) - 0.f) * A.dz));
            theta2 = lcon * sinf(6.28 * lcon_1 * (((x2 + 1) - 0.f) * A.dx));
            theta6 = lcon * sinf(6.28 * lcon_1 * (((z2 + 1) - 0.f) * A.dz));
            const float xx = ((x + 1) - 0.f) * A.dx, zz = ((z + 1) - 0.f) * A.dz;
            theta3 = lcon * sinf(6.28 * lcon_1 * xx);
            theta7 = lcon * sinf(6.28 * lcon_1 * zz);
            theta4 = lcon * sinf(6.28 * lcon_1 * xx);
            theta8 = lcon * sinf(6.28 * lcon_1 * zz);
        }
        float per_1, per_2, per_3, per_4, per_5, per_6;
        if (A.uniform_type == 0) per_1 = per_2 = per_3 = per_4 = per_5 = per_6 = A.const_period;
        else if (A.uniform_type == 1) { per_1 = per_2 = A.x_period; per_3 = per_4 = A.y_period; per_5 = per_6 = A.z_period; }
        else {
            const int sl = A.nx * A.ny;
            per_1 = period[x1 + y * A.nx + z * sl]; per_2 = period[x2 + y * A.nx + z * sl];
            per_3 = period[x + y1 * A.nx + z * sl]; per_4 = period[x + y2 * A.nx + z * sl];
            per_5 = period[x + y * A.nx + z1 * sl]; per_6 = period[x + y * A.nx + z2 * sl];
        }
        // i*cos(t) - j*sin(t)  and  i*sin(t) + j*cos(t): again first product rounded, second fused
        auto rot_x = [&](float t) { return __fmaf_rn(-fj, sinf(t), __fmul_rn(fi, cosf(t))); };
        auto rot_y = [&](float t) { return __fmaf_rn(fj, cosf(t), __fmul_rn(fi, sinf(t))); };
        float kx1 = k_scaled(per_1, rot_x(theta1)), kx2 = k_scaled(per_2, rot_x(theta2));
        float ky1 = k_scaled(per_3, rot_y(theta3)), ky2 = k_scaled(per_4, rot_y(theta4));
        float kz1 = k_scaled_int(per_5, f.z), kz2 = k_scaled_int(per_6, f.z);
        float phii = __fadd_rn(__fadd_rn(pair_sum(a1, kx1, a2, kx2), pair_sum(b1, ky1, b2, ky2)), pair_sum(c1, kz1, c2, kz2));
        if (A.latticetype == 's' && A.sinewave_zaxis) {
            const float fk = (float)f.z;
            kz1 = k_scaled(per_5, __fmaf_rn(fk, cosf(con * theta5 - angl), __fmul_rn(fj, sinf(theta5))));
            kz2 = k_scaled(per_6, __fmaf_rn(fk, cosf(con * theta6 - angl), __fmul_rn(fj, sinf(theta6))));
            ky1 = k_scaled(per_3, __fmaf_rn(-fk, sinf(con * theta7 + angl), __fmul_rn(fj, cosf(theta7))));
            ky2 = k_scaled(per_4, __fmaf_rn(-fk, sinf(con * theta8 + angl), __fmul_rn(fj, cosf(theta8))));
            kx1 = k_scaled_int(per_1, f.x);
            kx2 = k_scaled_int(per_2, f.x);
            phii = __fadd_rn(phii, __fadd_rn(__fadd_rn(pair_sum(a1, kx1, a2, kx2), pair_sum(b1, ky1, b2, ky2)), pair_sum(c1, kz1, c2, kz2)));
        }
        phi[tx] = phii;
    }
}

// ---------------------------------------------------------------- CG (GPUCG_lattice), all harmonics at once
// Per-harmonic state on the device.  `run` is 1 while the reference's `while (iCounter < iter && delta_new > term)` holds.
struct CgState { float delta_new, delta_old, temp, alpha, beta, res_best; int counter, run; };

// The reference reduces 1024 values with a shared-memory halving tree (GPUScalar_lattice_kernel :620-650: strides 512 ... 1, ten
// barriers).  The same additions, in the same association, with ONE barrier: after the strides 512 ... 32 element t (< 32) holds
// the tree sum of {c[t + 32 k]}, k paired as (k, k+16), (k, k+8), ..., which lane t of the first warp can form by itself from
// 32 conflict-free loads; the strides 16 ... 1 are shuffles.  Valid in lane 0 of warp 0 (returned to every lane of warp 0; other
// warps get an unspecified value).
__device__ __forceinline__ float block_tree_sum_1024(float c, float* cc) {
    const int tx = threadIdx.x;
    cc[tx] = c;
    __syncthreads();
    float r = 0.0f;
    if (tx < 32) {
        float a[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) a[k] = cc[tx + 32 * k];
#pragma unroll
        for (int half = 16; half > 0; half >>= 1)
#pragma unroll
            for (int k = 0; k < half; ++k) a[k] = __fadd_rn(a[k], a[k + half]);
        r = a[0];
#pragma unroll
        for (int s2 = 16; s2 > 0; s2 >>= 1) r = __fadd_rn(r, __shfl_down_sync(0xffffffffu, r, s2));
    }
    return r;
}
// Reduction_lattice (:22-74) on one harmonic's block partials: thread t sums partial[t], partial[t+1024], ... in order, then the tree
__device__ __forceinline__ float second_level_sum(const float* partial, int block_num, float* cc) {
    float c = 0.0f;
    for (int idx = threadIdx.x; idx < block_num; idx += 1024) c = __fadd_rn(c, partial[idx]);
    return block_tree_sum_1024(c, cc);
}

// 7-band stencil row of D^T D (GPUMatvec_lattice_kernel :420-597): returns diag contribution and the two off-diagonal terms
__device__ __forceinline__ void stencil_axis(const float* d, int tx, int v, int n, int stride, float& diag, float& t2, float& t3) {
    float p2, p3;
    if (v == 0) { p2 = d[tx + stride]; p3 = d[tx + 2 * stride]; diag = 1.25f; t2 = -p2; t3 = __fmul_rn(p3, -0.25f); }
    else if (v == 1) { p2 = d[tx - stride]; p3 = d[tx + 2 * stride]; diag = 1.25f; t2 = -p2; t3 = __fmul_rn(p3, -0.25f); }
    else if (v == n - 2) { p2 = d[tx - 2 * stride]; p3 = d[tx + stride]; diag = 1.25f; t2 = __fmul_rn(p2, -0.25f); t3 = -p3; }
    else if (v == n - 1) { p2 = d[tx - 2 * stride]; p3 = d[tx - stride]; diag = 1.25f; t2 = __fmul_rn(p2, -0.25f); t3 = -p3; }
    else { p2 = d[tx - 2 * stride]; p3 = d[tx + 2 * stride]; diag = 0.5f; t2 = __fmul_rn(p2, -0.25f); t3 = __fmul_rn(p3, -0.25f); }
}

// stage 0 (once): d = res = b, phi = 0, partial <res, res> and <res, d>
__global__ void __launch_bounds__(1024) cg_init_kernel(float* __restrict__ phi_all, float* __restrict__ d_all, float* __restrict__ res_all, float* __restrict__ partial,
                                                       int n, int block_num) {
    __shared__ float cc[1024];
    const int h = blockIdx.y, ind = blockIdx.x * 1024 + threadIdx.x;
    float c = 0.0f;
    if (ind < n) {
        const size_t o = (size_t)h * n + ind;
        const float b = phi_all[o];
        d_all[o] = b;
        res_all[o] = b;
        phi_all[o] = 0.0f;
        c = __fmul_rn(b, b);
    }
    const float s = block_tree_sum_1024(c, cc);
    if (threadIdx.x == 0) partial[(size_t)h * block_num + blockIdx.x] = s;
}
__global__ void __launch_bounds__(1024) cg_init_reduce_kernel(const float* __restrict__ partial, int block_num, CgState* st, int iter, float term) {
    __shared__ float cc[1024];
    const int h = blockIdx.x;
    const float s = second_level_sum(partial + (size_t)h * block_num, block_num, cc);
    if (threadIdx.x == 0) {
        CgState t;
        t.res_best = sqrtf(s);   // g_ResBest (unused afterwards, kept for completeness)
        t.delta_new = s;         // <res, d> with d == res: the same products in the same tree
        t.delta_old = s; t.temp = 0.f; t.alpha = 0.f; t.beta = 0.f;
        t.counter = 1;
        t.run = (t.counter < iter && t.delta_new > term) ? 1 : 0;
        st[h] = t;
    }
}
// stage 1: q = A d, partial <d, q>
__global__ void __launch_bounds__(1024) cg_matvec_kernel(const float* __restrict__ d_all, float* __restrict__ q_all, float* __restrict__ partial, const CgState* __restrict__ st,
                                                         int nx, int ny, int nz, int block_num, const Grid3 g3) {
    __shared__ float cc[1024];
    const int h = blockIdx.y;
    if (!st[h].run) return;
    const int n = nx * ny * nz, tx = blockIdx.x * 1024 + threadIdx.x;
    const float* d = d_all + (size_t)h * n;
    float c = 0.0f;
    if (tx < n) {
        int x, y, z;
        point_xyz((size_t)tx, g3, x, y, z);   // multiply-shift instead of two integer divisions per point
        const float phi1 = d[tx];
        float x1, x2, x3, y1, y2, y3, z1, z2, z3;
        stencil_axis(d, tx, x, nx, 1, x1, x2, x3);
        stencil_axis(d, tx, y, ny, nx, y1, y2, y3);
        stencil_axis(d, tx, z, nz, nx * ny, z1, z2, z3);
        // reference SASS: FADD, FADD, FFMA(phi1, s, x2), then five FADDs in source order
        float a = __fmaf_rn(phi1, __fadd_rn(__fadd_rn(x1, y1), z1), x2);
        a = __fadd_rn(a, x3); a = __fadd_rn(a, y2); a = __fadd_rn(a, y3); a = __fadd_rn(a, z2); a = __fadd_rn(a, z3);
        q_all[(size_t)h * n + tx] = a;
        c = __fmul_rn(phi1, a);
    }
    const float s = block_tree_sum_1024(c, cc);
    if (threadIdx.x == 0) partial[(size_t)h * block_num + blockIdx.x] = s;
}

// ---- float4 variants (nx % 4 == 0): a block is 256 threads x 4 consecutive points = the same 1024-point range, so the block
// partial goes through the same tree; a thread loads its row segment and the neighbouring segments as float4 (7 vector loads per
// 4 points in the interior instead of 13 scalar loads per point)
__device__ __forceinline__ float block_tree_sum_1024_from4(const float c[4], float* cc) {
    *reinterpret_cast<float4*>(cc + 4 * threadIdx.x) = make_float4(c[0], c[1], c[2], c[3]);
    __syncthreads();
    float r = 0.0f;
    if (threadIdx.x < 32) {
        float a[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) a[k] = cc[threadIdx.x + 32 * k];
#pragma unroll
        for (int half = 16; half > 0; half >>= 1)
#pragma unroll
            for (int k = 0; k < half; ++k) a[k] = __fadd_rn(a[k], a[k + half]);
        r = a[0];
#pragma unroll
        for (int s2 = 16; s2 > 0; s2 >>= 1) r = __fadd_rn(r, __shfl_down_sync(0xffffffffu, r, s2));
    }
    return r;
}
// the two off-diagonal terms of one axis for a whole float4 segment: which neighbour rows and which coefficients (Gratings.cu:
// 440-579); sa/sb = signed row offsets in units of `stride`
__device__ __forceinline__ void axis_case(int v, int n, int& oa, int& ob, bool& nega, bool& negb, float& diag) {
    if (v == 0) { oa = 1; ob = 2; nega = true; negb = false; diag = 1.25f; }
    else if (v == 1) { oa = -1; ob = 2; nega = true; negb = false; diag = 1.25f; }
    else if (v == n - 2) { oa = -2; ob = 1; nega = false; negb = true; diag = 1.25f; }
    else if (v == n - 1) { oa = -2; ob = -1; nega = false; negb = true; diag = 1.25f; }
    else { oa = -2; ob = 2; nega = false; negb = false; diag = 0.5f; }
}
__device__ __forceinline__ float off_term(float p, bool neg) { return neg ? -p : __fmul_rn(p, -0.25f); }

__global__ void __launch_bounds__(256) cg_matvec4_kernel(const float* __restrict__ d_all, float* __restrict__ q_all, float* __restrict__ partial,
                                                         const CgState* __restrict__ st, int nx, int ny, int nz, int block_num, const Grid3 g3) {
    __shared__ __align__(16) float cc[1024];
    const int h = blockIdx.y;
    if (!st[h].run) return;
    const int n = nx * ny * nz, e0 = blockIdx.x * 1024 + threadIdx.x * 4;
    const float* d = d_all + (size_t)h * n;
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    if (e0 < n) {
        int x0, y, z;
        point_xyz((size_t)e0, g3, x0, y, z);
        const float4 own = *reinterpret_cast<const float4*>(d + e0);
        const float4 lf = x0 >= 4 ? *reinterpret_cast<const float4*>(d + e0 - 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 rt = x0 + 4 < nx ? *reinterpret_cast<const float4*>(d + e0 + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float w[12] = {lf.x, lf.y, lf.z, lf.w, own.x, own.y, own.z, own.w, rt.x, rt.y, rt.z, rt.w};
        int oa, ob;
        bool nega, negb;
        float y1, z1;
        axis_case(y, ny, oa, ob, nega, negb, y1);
        const float4 ya = *reinterpret_cast<const float4*>(d + e0 + oa * nx), yb = *reinterpret_cast<const float4*>(d + e0 + ob * nx);
        const float y2[4] = {off_term(ya.x, nega), off_term(ya.y, nega), off_term(ya.z, nega), off_term(ya.w, nega)};
        const float y3[4] = {off_term(yb.x, negb), off_term(yb.y, negb), off_term(yb.z, negb), off_term(yb.w, negb)};
        const int sl = nx * ny;
        axis_case(z, nz, oa, ob, nega, negb, z1);
        const float4 za = *reinterpret_cast<const float4*>(d + e0 + oa * sl), zb = *reinterpret_cast<const float4*>(d + e0 + ob * sl);
        const float z2[4] = {off_term(za.x, nega), off_term(za.y, nega), off_term(za.z, nega), off_term(za.w, nega)};
        const float z3[4] = {off_term(zb.x, negb), off_term(zb.y, negb), off_term(zb.z, negb), off_term(zb.w, negb)};
        float q[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int x = x0 + j;
            int xa, xb;
            bool nxa, nxb;
            float x1;
            axis_case(x, nx, xa, xb, nxa, nxb, x1);
            const float x2 = off_term(w[4 + j + xa], nxa), x3 = off_term(w[4 + j + xb], nxb);
            const float phi1 = w[4 + j];
            // reference SASS: FADD, FADD, FFMA(phi1, s, x2), then five FADDs in source order
            float a = __fmaf_rn(phi1, __fadd_rn(__fadd_rn(x1, y1), z1), x2);
            a = __fadd_rn(a, x3); a = __fadd_rn(a, y2[j]); a = __fadd_rn(a, y3[j]); a = __fadd_rn(a, z2[j]); a = __fadd_rn(a, z3[j]);
            q[j] = a;
            c[j] = __fmul_rn(phi1, a);
        }
        *reinterpret_cast<float4*>(q_all + (size_t)h * n + e0) = make_float4(q[0], q[1], q[2], q[3]);
    }
    const float s = block_tree_sum_1024_from4(c, cc);
    if (threadIdx.x == 0) partial[(size_t)h * block_num + blockIdx.x] = s;
}
__global__ void __launch_bounds__(256) cg_update4_kernel(float* __restrict__ phi_all, float* __restrict__ res_all, const float* __restrict__ d_all,
                                                         const float* __restrict__ q_all, float* __restrict__ partial, const CgState* __restrict__ st, int n, int block_num) {
    __shared__ __align__(16) float cc[1024];
    const int h = blockIdx.y;
    if (!st[h].run) return;
    const int e0 = blockIdx.x * 1024 + threadIdx.x * 4;
    const float alpha = st[h].alpha, nalpha = (float)(-1.0 * (double)alpha);
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    if (e0 < n) {
        const size_t o = (size_t)h * n + e0;
        const float4 p = *reinterpret_cast<const float4*>(phi_all + o), r = *reinterpret_cast<const float4*>(res_all + o);
        const float4 dd = *reinterpret_cast<const float4*>(d_all + o), qq = *reinterpret_cast<const float4*>(q_all + o);
        const float pv[4] = {p.x, p.y, p.z, p.w}, rv[4] = {r.x, r.y, r.z, r.w}, dv[4] = {dd.x, dd.y, dd.z, dd.w}, qv[4] = {qq.x, qq.y, qq.z, qq.w};
        float po[4], ro[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            po[j] = __fmaf_rn(pv[j], 1.0f, __fmul_rn(alpha, dv[j]));
            ro[j] = __fmaf_rn(rv[j], 1.0f, __fmul_rn(nalpha, qv[j]));
            c[j] = __fmul_rn(ro[j], ro[j]);
        }
        *reinterpret_cast<float4*>(phi_all + o) = make_float4(po[0], po[1], po[2], po[3]);
        *reinterpret_cast<float4*>(res_all + o) = make_float4(ro[0], ro[1], ro[2], ro[3]);
    }
    const float s = block_tree_sum_1024_from4(c, cc);
    if (threadIdx.x == 0) partial[(size_t)h * block_num + blockIdx.x] = s;
}
__global__ void __launch_bounds__(256) cg_direction4_kernel(float* __restrict__ d_all, const float* __restrict__ res_all, const CgState* __restrict__ st, int n) {
    const int h = blockIdx.y;
    if (!st[h].run) return;
    const int e0 = blockIdx.x * 1024 + threadIdx.x * 4;
    const float beta = st[h].beta;
    if (e0 < n) {
        const size_t o = (size_t)h * n + e0;
        const float4 dd = *reinterpret_cast<const float4*>(d_all + o), r = *reinterpret_cast<const float4*>(res_all + o);
        *reinterpret_cast<float4*>(d_all + o) = make_float4(__fmaf_rn(dd.x, beta, __fmul_rn(1.0f, r.x)), __fmaf_rn(dd.y, beta, __fmul_rn(1.0f, r.y)),
                                                            __fmaf_rn(dd.z, beta, __fmul_rn(1.0f, r.z)), __fmaf_rn(dd.w, beta, __fmul_rn(1.0f, r.w)));
    }
}

// stage 2: temp = <d, q>, alpha = delta_new / temp
__global__ void __launch_bounds__(1024) cg_alpha_kernel(const float* __restrict__ partial, int block_num, CgState* st) {
    __shared__ float cc[1024];
    const int h = blockIdx.x;
    if (!st[h].run) return;
    const float s = second_level_sum(partial + (size_t)h * block_num, block_num, cc);
    if (threadIdx.x == 0) { st[h].temp = s; st[h].alpha = __fdiv_rn(st[h].delta_new, s); }
}
// stage 3: phi += alpha d ; res -= alpha q ; partial <res, res>      (VecSMultAddKernel_lattice: fma(V, a1, a2 * W))
__global__ void __launch_bounds__(1024) cg_update_kernel(float* __restrict__ phi_all, float* __restrict__ res_all, const float* __restrict__ d_all,
                                                         const float* __restrict__ q_all, float* __restrict__ partial, const CgState* __restrict__ st, int n, int block_num) {
    __shared__ float cc[1024];
    const int h = blockIdx.y;
    if (!st[h].run) return;
    const int ind = blockIdx.x * 1024 + threadIdx.x;
    const float alpha = st[h].alpha, nalpha = (float)(-1.0 * (double)alpha);
    float c = 0.0f;
    if (ind < n) {
        const size_t o = (size_t)h * n + ind;
        phi_all[o] = __fmaf_rn(phi_all[o], 1.0f, __fmul_rn(alpha, d_all[o]));
        const float r = __fmaf_rn(res_all[o], 1.0f, __fmul_rn(nalpha, q_all[o]));
        res_all[o] = r;
        c = __fmul_rn(r, r);
    }
    const float s = block_tree_sum_1024(c, cc);
    if (threadIdx.x == 0) partial[(size_t)h * block_num + blockIdx.x] = s;
}
// stage 4: delta_old = delta_new, delta_new = <res, res>, beta = delta_new / delta_old; the loop condition for the NEXT iteration
__global__ void __launch_bounds__(1024) cg_beta_kernel(const float* __restrict__ partial, int block_num, CgState* st, int iter, float term, int* running) {
    __shared__ float cc[1024];
    const int h = blockIdx.x;
    if (!st[h].run) return;
    const float s = second_level_sum(partial + (size_t)h * block_num, block_num, cc);
    if (threadIdx.x == 0) {
        CgState t = st[h];
        t.delta_old = t.delta_new;
        t.delta_new = s;
        t.beta = __fdiv_rn(t.delta_new, t.delta_old);
        st[h] = t;
    }
}
// stage 5: d = beta d + res (always executed for a running harmonic, as in the reference), then counter++ and the loop test
__global__ void __launch_bounds__(1024) cg_direction_kernel(float* __restrict__ d_all, const float* __restrict__ res_all, CgState* st, int n, int iter, float term, int* running) {
    const int h = blockIdx.y;
    if (!st[h].run) return;
    const int ind = blockIdx.x * 1024 + threadIdx.x;
    const float beta = st[h].beta;
    if (ind < n) {
        const size_t o = (size_t)h * n + ind;
        d_all[o] = __fmaf_rn(d_all[o], beta, __fmul_rn(1.0f, res_all[o]));
    }
}
__global__ void cg_advance_kernel(CgState* st, int nharm, int iter, float term, int* running) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nharm || !st[h].run) return;
    const int cnt = st[h].counter + 1;
    st[h].counter = cnt;
    if (!(cnt < iter && st[h].delta_new > term)) { st[h].run = 0; atomicSub(running, 1); }
}
__global__ void cg_count_running_kernel(const CgState* st, int nharm, int* running) {
    int r = 0;
    for (int h = 0; h < nharm; ++h) r += st[h].run;
    *running = r;
}

} // namespace

int k_finding_phi(Ctx* c, float* phi_all, const float* period, const int* ijk_host, int nharm, int nx, int ny, int nz, float dx, float dy, float dz, int latticetype,
                  int uniform_type, float const_period, float x_period, float y_period, float z_period, float lcon, float lcon_1, int sinewave_zaxis) {
    if (nharm <= 0 || nx < 4 || ny < 4 || nz < 4) return fail_msg(c, "finding_phi: needs at least one harmonic and 4 points per axis");
    if (uniform_type == 2 && !period) return fail_msg(c, "finding_phi: period field missing");
    if ((size_t)nx * ny * nz > 0x7fffffffu) return fail_msg(c, "finding_phi: control grid too large");
    int3* d_ijk;
    GCB_CHECK(c, cudaMalloc(&d_ijk, (size_t)nharm * sizeof(int3)));
    cudaError_t e = cudaMemcpyAsync(d_ijk, ijk_host, (size_t)nharm * sizeof(int3), cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) { cudaFree(d_ijk); return fail(c, "finding_phi: harmonic list", e); }
    PhiArgs A{nx, ny, nz, dx, dy, dz, latticetype, uniform_type, const_period, x_period, y_period, z_period, lcon, lcon_1, sinewave_zaxis};
    const int n = nx * ny * nz;
    dim3 grid(std::min((unsigned)((n + 255) / 256), (unsigned)c->num_sms * 8), nharm);
    finding_phi_kernel<<<grid, 256, 0, c->stream>>>(phi_all, period, d_ijk, nharm, A);
    c->launches++;
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_ijk);
    if (e != cudaSuccess) return fail(c, "finding_phi", e);
    return 0;
}

int k_cg_batched(Ctx* c, float* phi_all, int nharm, int nx, int ny, int nz, int iter, float end_res, int* final_iter, float* final_res) {
    if (nharm <= 0) return 0;
    if (nx < 4 || ny < 4 || nz < 4) return fail_msg(c, "GPUCG_lattice: needs at least 4 points per axis");
    const size_t n = (size_t)nx * ny * nz;
    if (n > 0x7fffffffu) return fail_msg(c, "GPUCG_lattice: control grid too large");
    const int block_num = (int)((n + 1023) / 1024);
    const float term = end_res * end_res;
    float *d_d = nullptr, *d_q = nullptr, *d_res = nullptr, *partial = nullptr;
    CgState* st = nullptr;
    int* running = nullptr;
    auto cleanup = [&]() { cudaFree(d_d); cudaFree(d_q); cudaFree(d_res); cudaFree(partial); cudaFree(st); cudaFree(running); };
#define CG_CHECK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return fail(c, "GPUCG_lattice: " #call, e_); } } while (0)
    CG_CHECK(cudaMalloc(&d_d, n * nharm * sizeof(float)));
    CG_CHECK(cudaMalloc(&d_q, n * nharm * sizeof(float)));
    CG_CHECK(cudaMalloc(&d_res, n * nharm * sizeof(float)));
    CG_CHECK(cudaMalloc(&partial, (size_t)block_num * nharm * sizeof(float)));
    CG_CHECK(cudaMalloc(&st, (size_t)nharm * sizeof(CgState)));
    CG_CHECK(cudaMalloc(&running, sizeof(int)));
    cudaStream_t s = c->stream;
    const dim3 grid(block_num, nharm);
    const Grid3 g3 = make_grid3((size_t)nx, (size_t)ny, (size_t)nz);
    const bool vec4 = (nx % 4 == 0) && (((uintptr_t)phi_all & 15) == 0) && g3.small;  // rows are whole float4 groups
    cg_init_kernel<<<grid, 1024, 0, s>>>(phi_all, d_d, d_res, partial, (int)n, block_num);
    cg_init_reduce_kernel<<<nharm, 1024, 0, s>>>(partial, block_num, st, iter, term);
    cg_count_running_kernel<<<1, 1, 0, s>>>(st, nharm, running);
    c->launches += 3;
    int h_running = 1;
    const int poll = 8;  // iterations between two looks at the "still running" counter
    for (int it = 1; it < iter && h_running > 0; it += poll) {
        for (int k = 0; k < poll && it + k < iter; ++k) {
            if (vec4) cg_matvec4_kernel<<<grid, 256, 0, s>>>(d_d, d_q, partial, st, nx, ny, nz, block_num, g3);
            else cg_matvec_kernel<<<grid, 1024, 0, s>>>(d_d, d_q, partial, st, nx, ny, nz, block_num, g3);
            cg_alpha_kernel<<<nharm, 1024, 0, s>>>(partial, block_num, st);
            if (vec4) cg_update4_kernel<<<grid, 256, 0, s>>>(phi_all, d_res, d_d, d_q, partial, st, (int)n, block_num);
            else cg_update_kernel<<<grid, 1024, 0, s>>>(phi_all, d_res, d_d, d_q, partial, st, (int)n, block_num);
            cg_beta_kernel<<<nharm, 1024, 0, s>>>(partial, block_num, st, iter, term, running);
            if (vec4) cg_direction4_kernel<<<grid, 256, 0, s>>>(d_d, d_res, st, (int)n);
            else cg_direction_kernel<<<grid, 1024, 0, s>>>(d_d, d_res, st, (int)n, iter, term, running);
            cg_advance_kernel<<<(nharm + 63) / 64, 64, 0, s>>>(st, nharm, iter, term, running);
            c->launches += 6;
        }
        CG_CHECK(cudaMemcpyAsync(&h_running, running, sizeof(int), cudaMemcpyDeviceToHost, s));
        CG_CHECK(cudaStreamSynchronize(s));
    }
    std::vector<CgState> hs(nharm);
    CG_CHECK(cudaMemcpyAsync(hs.data(), st, (size_t)nharm * sizeof(CgState), cudaMemcpyDeviceToHost, s));
    CG_CHECK(cudaStreamSynchronize(s));
    CG_CHECK(cudaGetLastError());
    for (int h = 0; h < nharm; ++h) {
        if (final_iter) final_iter[h] = hs[h].counter;
        if (final_res) final_res[h] = sqrtf(hs[h].delta_new);  // FinalRes = sqrt(g_delta_new) (:966), host float sqrt
    }
#undef CG_CHECK
    cleanup();
    return 0;
}

} // namespace gcb
