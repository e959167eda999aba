// fdga_krylov.cuh -- device-resident vector kernels of DQGMRES for the mfRG linear map.
// The reference solves  mfRGLinearMap(S, strategy) * x = R_F  with Krylov.dqgmres on HOST vectors of length(S.F)
// (src/mfRG.jl:147-151): every matvec moves x and y across PCIe and the incomplete orthogonalisation runs on the CPU.
// Here the Krylov basis V, the direction vectors P, the iterate x and the work vector w all live in HBM; per iteration
// only the (<= memory + 1) Hessenberg entries cross to the host, which keeps the Givens rotations in double precision.
#pragma once
#include "fdga_kernels.cuh"

namespace fdga {

#define FDGA_KRY_BLOCKS 592     // 148 SMs x 4 resident CTAs: one wave of grid-stride CTAs
#define FDGA_KRY_THREADS 256
#define FDGA_KRY_CHUNK 128      // direction vectors combined per launch of kry_direction_kernel

// One modified-Gram-Schmidt step, fused:  w -= h_sub * v_sub  (if v_sub), then  out = <v_dot, w> = sum conj(v_dot) w
// (v_dot == nullptr: <w, w>).  h_sub is a DEVICE scalar written by the previous step, so a whole orthogonalisation
// sweep is queued without host synchronisation.  Deterministic: per-CTA partial sums, the last CTA (ticket) adds
// them in index order.
__global__ void __launch_bounds__(FDGA_KRY_THREADS)
kry_mgs_kernel(C* __restrict__ w, const C* __restrict__ v_sub, const C* __restrict__ h_sub, const C* __restrict__ v_dot,
               C* __restrict__ out, C* __restrict__ part, unsigned int* __restrict__ ticket, long long n) {
    __shared__ bool is_last;
    C h = zeroC();
    if (v_sub != nullptr) h = *h_sub;
    C acc = zeroC();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        C wi = w[i];
        if (v_sub != nullptr) { wi = wi - h * v_sub[i]; w[i] = wi; }
        const C d = (v_dot != nullptr) ? v_dot[i] : wi;
        acc += conjC(d) * wi;
    }
    acc = block_reduce(acc);
    if (threadIdx.x == 0) {
        part[blockIdx.x] = acc;
        __threadfence();
        is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    C tot = zeroC();
    for (int j = threadIdx.x; j < (int)gridDim.x; j += blockDim.x) tot += part[j];     // fixed order per thread
    tot = block_reduce(tot);
    if (threadIdx.x == 0) { *out = tot; *ticket = 0u; }
}

// Blocked form of the same sweep: up to FDGA_KRY_B basis vectors per launch.  For the new block {v_j} the kernel accumulates
// d_j = <v_j, w'> (w' = w minus the previous block, subtracted on the fly) AND the Gram entries G_jl = <v_j, v_l>, l < j; the last
// CTA then runs the modified Gram-Schmidt recurrence on the coefficients,  h_j = d_j - sum_{l<j} h_l G_jl  (= <v_j, w' - sum_{l<j}
// h_l v_l> exactly), so the algebra is MGS while w is read and written once per block instead of once per vector:
// 2.5 vector passes per basis vector instead of 4, a quarter of the launches.
#define FDGA_KRY_B 4
struct KryBlk { int nprev, nnew, norm; int prev_slot[FDGA_KRY_B], new_slot[FDGA_KRY_B]; };
__global__ void __launch_bounds__(FDGA_KRY_THREADS)
kry_bmgs_kernel(C* __restrict__ w, const C* __restrict__ Vbase, long long n, const __grid_constant__ KryBlk blk,
                const C* __restrict__ h_prev, C* __restrict__ h_new, C* __restrict__ part, unsigned int* __restrict__ ticket) {
    constexpr int B = FDGA_KRY_B, NACC = B + B * (B - 1) / 2 + 1;
    __shared__ bool is_last;
    __shared__ C tot[NACC];
    C hp[B];
#pragma unroll
    for (int p = 0; p < B; ++p) hp[p] = p < blk.nprev ? h_prev[p] : zeroC();
    C acc[NACC];
#pragma unroll
    for (int a = 0; a < NACC; ++a) acc[a] = zeroC();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        C wi = w[i];
        if (blk.nprev > 0) {
#pragma unroll
            for (int p = 0; p < B; ++p) if (p < blk.nprev) wi = wi - hp[p] * Vbase[(size_t)blk.prev_slot[p] * n + i];
            w[i] = wi;
        }
        C v[B];
#pragma unroll
        for (int j = 0; j < B; ++j) v[j] = j < blk.nnew ? Vbase[(size_t)blk.new_slot[j] * n + i] : zeroC();
        int g = B;
#pragma unroll
        for (int j = 0; j < B; ++j) {
            acc[j] += conjC(v[j]) * wi;
#pragma unroll
            for (int l = 0; l < j; ++l) { acc[g] += conjC(v[j]) * v[l]; ++g; }
        }
        if (blk.norm) acc[NACC - 1] += conjC(wi) * wi;
    }
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
        C r = block_reduce(acc[a]);
        if (threadIdx.x == 0) part[(size_t)blockIdx.x * NACC + a] = r;
    }
    if (threadIdx.x == 0) { __threadfence(); is_last = atomicAdd(ticket, 1u) == gridDim.x - 1; }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
        C t = zeroC();
        for (int j = threadIdx.x; j < (int)gridDim.x; j += blockDim.x) t += part[(size_t)j * NACC + a];
        t = block_reduce(t);
        if (threadIdx.x == 0) tot[a] = t;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        C h[B];
        int g = B;
        for (int j = 0; j < B; ++j) {
            C x = tot[j];
            for (int l = 0; l < j; ++l) { x = x - h[l] * tot[g]; ++g; }
            h[j] = x;
            if (j < blk.nnew) h_new[j] = x;
        }
        if (blk.norm) h_new[blk.nnew] = tot[NACC - 1];
        *ticket = 0u;
    }
}

struct KryCoefs { int nc; int slot[FDGA_KRY_CHUNK]; C t[FDGA_KRY_CHUNK]; };
// p_new = (src - sum_j t_j P[slot_j]) ; on the last chunk: p_new *= inv_r and x += gamma * p_new
// (p_m = (v_m - sum_i r_{i,m} p_i) / r_{m,m},  x_m = x_{m-1} + gamma_m p_m)
__global__ void __launch_bounds__(FDGA_KRY_THREADS)
kry_direction_kernel(C* p_new, const C* src, const C* Pbase, long long n,     // p_new may alias one ring slot of Pbase
                     const __grid_constant__ KryCoefs cf, int last, C inv_r, C gamma, C* __restrict__ x) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        C p = src[i];
        for (int j = 0; j < cf.nc; ++j) p = p - cf.t[j] * Pbase[(size_t)cf.slot[j] * n + i];
        if (last) { p = p * inv_r; x[i] += gamma * p; }
        p_new[i] = p;
    }
}

}  // namespace fdga
