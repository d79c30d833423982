// Faithful micro-replica of the solver's per-group step, pieces switched on one by one.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint32_t u32; typedef uint64_t u64;
#define HQS_AMOUNT_MAX (~(u64)0)
constexpr int RT = 4;
struct VarT { u64 amount[RT]; float rcpf[2 * RT]; u64 min_time_ms; u32 all_mask; u32 used_mask; };
struct ClassT { u32 n_variants; u32 pad; VarT v[8]; };
struct GroupOut { u32 k, out_off, seg_lo, seg_n; };

__device__ __forceinline__ u64 fit_count(const u64 (&fr)[RT], const u64 (&tot)[RT], const VarT& dv, u64 cap) {
    u64 cnt = cap; bool big = false;
    const u32 used = dv.used_mask, allm = dv.all_mask;
#pragma unroll
    for (int r = 0; r < RT; ++r) {
        const bool on = (used >> r) & 1, all = (allm >> r) & 1;
        const u64 n = fr[r], d = dv.amount[r];
        const bool fits_cap = __umul64hi(d, cap) == 0 && d * cap <= n;
        const float nf = __fmaf_rn(__uint2float_rn((u32)(n >> 32)), 4294967296.0f, __uint2float_rn((u32)n));
        const float qf = nf * dv.rcpf[r];
        u64 q = (u64)__float2uint_rz(fminf(qf, 1048576.0f));
        u64 p = q * d;
        q = p > n ? q - 1 : (n - p >= d ? q + 1 : q);
        p = q * d;
        q = p > n ? q - 1 : (n - p >= d ? q + 1 : q);
        const u64 q_all = (tot[r] != 0 && n == tot[r]) ? 1 : 0;
        const bool unconstrained = !on || (!all && (n == HQS_AMOUNT_MAX || fits_cap));
        big |= on && !all && !unconstrained && qf >= 1048576.0f;
        const u64 qr = all ? q_all : q;
        cnt = unconstrained ? cnt : (cnt < qr ? cnt : qr);
    }
    if (big) { cnt = cap;
#pragma unroll
        for (int r = 0; r < RT; ++r) { if (!((used >> r) & 1)) continue; u64 q; if ((allm >> r) & 1) q = (tot[r] != 0 && fr[r] == tot[r]) ? 1 : 0; else if (fr[r] != HQS_AMOUNT_MAX) q = fr[r] / dv.amount[r]; else continue; cnt = cnt < q ? cnt : q; } }
    return cnt;
}

template <int MODE>
__global__ void k(const ClassT* g_classes, const uint2* g_glist, int n_list, u64* out, long long* cyc, GroupOut* g_gout) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ u64 s_x[64], s_f[64];
    ClassT* classes = reinterpret_cast<ClassT*>(smem);
    uint2* s_glist = reinterpret_cast<uint2*>(smem + 16 * sizeof(ClassT));
    u32* s_segc = reinterpret_cast<u32*>(s_glist + 256);
    u32* s_segw = s_segc + 4096;
    GroupOut* s_gout = reinterpret_cast<GroupOut*>(s_segw + 4096);
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    for (u32 i = tid; i < 16 * sizeof(ClassT) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(g_classes)[i];
    for (u32 i = tid; i < (u32)n_list; i += blockDim.x) s_glist[i] = g_glist[i];
    u64 fr[RT] = {1310720000ull, 81920000ull, 5242880000ull, 20971520000ull}, tot[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) tot[r] = fr[r];
    u32 parity = 0, seg_base = 0, out_base = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int e = 0; e < n_list; ++e) {
        const uint2 ge = s_glist[e];
        const u32 c = ge.x & 15, n_all = ge.y;
        u32 remaining = n_all;
        const u32 seg_lo = seg_base;
        const VarT& dv = classes[c].v[0];
        u64 cnt = 0;
        if (MODE >= 1) cnt = fit_count(fr, tot, dv, remaining); else cnt = (tid >= (u32)(e & 127)) ? remaining : 0;
        const bool can1 = cnt != 0, can_all = cnt >= remaining;
        u32 take = 0, exc_cnt = 0, seg_rank = 0, n_takers = 0, handed = 0;
        if (MODE >= 0) {
            u64* fb = s_f + 32 * (parity & 1); parity++;
            const u32 has = __ballot_sync(0xffffffffu, can1);
            const u32 first = has ? (u32)(__ffs(has) - 1) : 0u;
            const u32 fall = __shfl_sync(0xffffffffu, can_all ? 1u : 0u, first);
            if (lane == 0) fb[warp] = has ? (2ull | fall) : 0ull;
            __syncthreads();
            const u64 ee = lane < nwarps ? fb[lane] : 0ull;
            const u32 anyw = __ballot_sync(0xffffffffu, ee != 0);
            if (anyw) {
                const u32 wf = (u32)(__ffs(anyw) - 1);
                const u64 ef = __shfl_sync(0xffffffffu, ee, wf);
                if ((ef & 1ull) && MODE != 3) { take = (warp == wf && lane == first && can1) ? remaining : 0; n_takers = 1; handed = remaining; }
                else {
                    u64* buf = s_x + 32 * (parity & 1); parity++;
                    u64 inc = cnt;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { const u64 y = __shfl_up_sync(0xffffffffu, inc, d); if ((int)lane >= d) inc += y; }
                    const u32 hasb = __ballot_sync(0xffffffffu, cnt != 0);
                    if (lane == 31) buf[warp] = inc | ((u64)__popc(hasb) << 42);
                    __syncthreads();
                    u64 below = 0, all = 0;
                    for (u32 w2 = 0; w2 < nwarps; ++w2) { const u64 v = buf[w2]; all += v; if (w2 < warp) below += v; }
                    const u64 mask = (1ull << 42) - 1;
                    const u64 exc = (below & mask) + inc - cnt;
                    exc_cnt = (u32)(exc < remaining ? exc : remaining);
                    if (cnt && exc < remaining) { const u64 room = remaining - exc; take = (u32)(cnt < room ? cnt : room); }
                    seg_rank = (u32)(below >> 42) + __popc(hasb & ((1u << lane) - 1));
                    n_takers = (u32)__syncthreads_count(take != 0);
                    const u64 tc = all & mask;
                    handed = (u32)(tc < remaining ? tc : remaining);
                }
            }
        } else { take = (u32)cnt; handed = remaining; n_takers = 1; }
        if (MODE >= 4 && take) {
            const u32 si = seg_base + seg_rank;
            s_segc[si & 4095] = (n_all - remaining) + exc_cnt + take;
            s_segw[si & 4095] = tid;
#pragma unroll
            for (int r = 0; r < RT; ++r) if ((dv.used_mask >> r) & 1) fr[r] -= (u64)take * dv.amount[r];
        }
        remaining -= handed; seg_base += n_takers;
        if (MODE >= 5 && tid == 0) { GroupOut go; go.k = n_all - remaining; go.out_off = out_base; go.seg_lo = seg_lo; go.seg_n = seg_base - seg_lo; s_gout[e] = go; }
        out_base += n_all - remaining;
    }
    const long long t1 = clock64();
    if (tid == 0) *cyc = t1 - t0;
    out[tid] = fr[0] + fr[1] + seg_base + out_base;
    if (tid < (u32)n_list) g_gout[tid] = s_gout[tid];
}

int main() {
    const int n_list = 128;
    ClassT h_cls[16]; uint2 h_gl[128];
    for (int c = 0; c < 16; ++c) {
        h_cls[c].n_variants = 1;
        VarT& v = h_cls[c].v[0];
        u64 am[4] = {(u64)(1 + c % 16) * 10000, (u64)(c % 5) * 2500, (u64)(1 + (c * 7) % 64) * 10000, (u64)((c * 3) % 33) * 10000};
        v.used_mask = 0; v.all_mask = 0; v.min_time_ms = 0;
        for (int r = 0; r < 4; ++r) { v.amount[r] = am[r]; v.rcpf[r] = am[r] ? 1.0f / (float)am[r] : 0.f; if (am[r]) v.used_mask |= 1u << r; }
    }
    for (int e = 0; e < 128; ++e) h_gl[e] = make_uint2((e / 16) * 16 + (e % 16), 7800 + 13 * e);
    ClassT* d_cls; uint2* d_gl; u64* d_out; long long* d_cyc; GroupOut* d_go;
    cudaMalloc(&d_cls, sizeof h_cls); cudaMalloc(&d_gl, sizeof h_gl); cudaMalloc(&d_out, 1024 * 8); cudaMalloc(&d_cyc, 8); cudaMalloc(&d_go, 256 * 16);
    cudaMemcpy(d_cls, h_cls, sizeof h_cls, cudaMemcpyHostToDevice); cudaMemcpy(d_gl, h_gl, sizeof h_gl, cudaMemcpyHostToDevice);
    const size_t smem = 16 * sizeof(ClassT) + 256 * 8 + 2 * 4096 * 4 + 256 * 16;
    const char* names[] = {"exchange + fast path (trivial count)", "fit_count + exchange + fast path", "same", "  scan forced on every step", "+ segments, take (loop-carried free)", "+ group record"};
#define RUN(M) { cudaFuncSetAttribute(k<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); for (int rep = 0; rep < 2; ++rep) { k<M><<<1, 256, smem>>>(d_cls, d_gl, n_list, d_out, d_cyc, d_go); cudaDeviceSynchronize(); } long long c; cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost); printf("%-45s %8.1f cycles/group\n", names[M], (double)c / n_list); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
