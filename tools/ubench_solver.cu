// Micro-benchmark of the primitives on the solver's sequential critical path (one CTA, 256 threads).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench tools/ubench_solver.cu && /tmp/ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint32_t u32; typedef uint64_t u64;

__device__ __forceinline__ u64 div_floor(u64 n, u64 d, double rcp) {
    if (n >> 53) return n / d;
    u64 q = (u64)__double2ull_rz(__dmul_rn(__ull2double_rn(n), rcp));
    const u64 p = q * d;
    if (p > n) --q; else if (n - p >= d) ++q;
    return q;
}

template <int MODE>
__global__ void k(u64* out, long long* cyc, int iters, u64 seed, double rcp0) {
    __shared__ u64 s_x[64];
    __shared__ u64 s_f[64];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    u64 fr[4] = {seed + threadIdx.x * 7919ull, seed * 3 + threadIdx.x, seed * 5 + 11, seed * 7 + 13};
    u64 am[4] = {80000, 2500, 130000, 90000};
    double rc[4] = {rcp0 / 80000.0, rcp0 / 2500.0, rcp0 / 130000.0, rcp0 / 90000.0};
    u32 parity = 0;
    u64 acc = 0;
    u32 remaining = 8000;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        u64 cnt = 1;
        if (MODE >= 3) {          // exact fit count: 4 reciprocal divisions
            cnt = remaining;
#pragma unroll
            for (int r = 0; r < 4; ++r) { const u64 q = div_floor(fr[r], am[r], rc[r]); cnt = cnt < q ? cnt : q; }
        }
        if (MODE == 1 || MODE >= 4) {   // exchange 1: ballot + smem + barrier + read
            u64* fb = s_f + 32 * (parity & 1);
            const u32 has = __ballot_sync(0xffffffffu, cnt != 0);
            if (lane == 0) fb[warp] = has;
            __syncthreads();
            const u64 ee = lane < nwarps ? fb[lane] : 0ull;
            acc += __ballot_sync(0xffffffffu, ee != 0);
        }
        if (MODE == 2 || MODE >= 4) {   // block scan (u64 packed) + count barrier
            u64* buf = s_x + 32 * (parity & 1);
            const u64 x = cnt | (cnt ? (1ull << 42) : 0ull);
            u64 inc = x;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const u64 y = __shfl_up_sync(0xffffffffu, inc, d); if ((int)lane >= d) inc += y; }
            if (lane == 31) buf[warp] = inc;
            __syncthreads();
            u64 winc = lane < nwarps ? buf[lane] : 0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const u64 y = __shfl_up_sync(0xffffffffu, winc, d); if ((int)lane >= d) winc += y; }
            const u64 bt = __shfl_sync(0xffffffffu, winc, 31);
            const u64 wo = warp ? __shfl_sync(0xffffffffu, winc, warp - 1) : 0;
            const u64 exc = wo + inc - x;
            u32 take = 0;
            if (cnt && (exc & ((1ull << 42) - 1)) < remaining) take = 1;
            const u32 nt = __syncthreads_count(take != 0);
            acc += bt + nt;
            if (take) fr[0] -= am[0];
        }
        if (MODE == 0) { __syncthreads(); __syncthreads(); __syncthreads(); }
        parity++;
        remaining = 8000 + (u32)(acc & 1);
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) *cyc = t1 - t0;
    out[threadIdx.x] = acc + fr[0] + fr[1];
}

int main() {
    u64* d_out; long long* d_cyc; cudaMalloc(&d_out, 1024 * 8); cudaMalloc(&d_cyc, 8);
    const int iters = 2000;
    const char* names[] = {"3 barriers", "exchange (ballot+STS+BAR+LDS)", "block scan u64 + count barrier", "4 reciprocal divisions", "exchange + division + scan (general step)"};
    for (int mode = 0; mode < 5; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            if (mode == 0) k<0><<<1, 256>>>(d_out, d_cyc, iters, 1300000000ull, 1.0);
            if (mode == 1) k<1><<<1, 256>>>(d_out, d_cyc, iters, 1300000000ull, 1.0);
            if (mode == 2) k<2><<<1, 256>>>(d_out, d_cyc, iters, 1300000000ull, 1.0);
            if (mode == 3) k<3><<<1, 256>>>(d_out, d_cyc, iters, 1300000000ull, 1.0);
            if (mode == 4) k<4><<<1, 256>>>(d_out, d_cyc, iters, 1300000000ull, 1.0);
            cudaDeviceSynchronize();
        }
        long long c; cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-45s %8.1f cycles/iter\n", names[mode], (double)c / iters);
    }
    return 0;
}
