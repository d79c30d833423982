// hqs_solver.cuh — solve_k: the sequential first-fit solver (CTA 0), the chunk-table scan and the pack warps, plus
// xchg_k / the system-scope flag helpers of the sharded tick.  Included by hqsched.cu inside its anonymous namespace.
#pragma once

// ------------------------------------------------------------------------------------------------
// K2: solve_k (cooperative launch).
//   CTA 0          sequential priority-ordered first-fit, one thread per worker (solve_body)
//   CTAs 1..S      exclusive scan of table[][g] over chunks, one warp per group column
//   CTAs >= 1      then wait for CTA 0: if the tick has a saturated level, every worker is filled by its
//                  own warp (pack_body, alignment heuristic) spread over the whole grid
// ------------------------------------------------------------------------------------------------
constexpr u32 SEG_SMEM = 4096;        // count segments buffered in shared memory before a bulk flush
constexpr u32 PACK_MAX_CAND = 64;    // (class, variant) candidates of the packed level: 2 per lane
constexpr u32 PACK_MAX_ITER = 64;
constexpr u32 PACK_CHUNK_DIV = 8;
constexpr u32 PHASE_WAIT = 0, PHASE_PACK = 1, PHASE_EXIT = 2;
constexpr u32 HQS_MAX_PEERS = 16;      // ranks of a sharded ready set
constexpr long long SPIN_TIMEOUT_CYCLES = 4000000000ll;   // ~2 s: a stuck grid must not hang the GPU

// Amounts come in two widths.  u64: the ABI's fixed-point fractions as they are.  u32 ("narrow"): the same
// amounts divided by the per-resource gcd of all requested amounts — fit counts are unchanged by that
// (floor(n / d) == floor(floor(n / g) / (d / g)) when g divides d), the quotient estimate needs one int->float
// conversion and one fix-up instead of a 64-bit sequence, and the solver's sequential critical path shrinks
// accordingly.  The narrow path is taken when every scaled amount of the tick is below 2^31.
template <int RT, typename AT = u64>
struct VarT {
    AT amount[RT];
    float rcpf[2 * RT];  // [0, RT): fp32 1.0 / amount (0 where unused); [RT, 2 RT): fp32 of the exact amount
    u64 min_time_ms;
    u32 all_mask;
    u32 used_mask;
};
template <int RT, typename AT = u64>
struct ClassT {
    u32 n_variants;
    u32 pad;
    VarT<RT, AT> v[HQS_MAX_VARIANTS];
};
template <typename AT> struct AmountMax;
template <> struct AmountMax<u64> { static constexpr u64 value = HQS_AMOUNT_MAX; };
template <> struct AmountMax<u32> { static constexpr u32 value = 0xFFFFFFFFu; };
constexpr u64 NARROW_LIMIT = 0x7FFFFFFFull;      // scaled amounts of the narrow path stay below 2^31

struct SolveSync {
    u32 phase;
    u32 done;
};

struct PackScratch {          // global memory, written by CTA 0, read by the pack warps (and back)
    u64* fr;                  // [W][R]
    u32* quota;               // [W][PACK_MAX_CAND]   per (worker, group of the level)
    u32* taken;               // [W][PACK_MAX_CAND]   per (worker, candidate)
    u32* cand;                // [PACK_MAX_CAND]      class | variant << 16 | group-in-level << 24
    u32* meta;                // [2] n_cand, n_groups
};

struct SolveArgs {
    // tick input (device copy of the host staging buffer)
    const u64* free_rw;      // [W][R]
    const u64* total_rw;     // [W][R]
    const u64* rem_time;     // [W]
    const u32* order;        // [Q] class ids in processing order inside one priority level
    const uint8_t* vorder;   // [Q][HQS_MAX_VARIANTS] variant ids in first-fit order
    const uint8_t* blocked;  // [W][Q] bytes (bit v) or nullptr
    const void* classes;     // ClassT<RT, AT>[Q] of the solver's width
    const void* classes64;   // ClassT<RT, u64>[Q] (the pack warps work on exact amounts)
    u64 gscale[HQS_MAX_RESOURCES];   // narrow path: amount = scaled amount * gscale[r] (+ a per-worker remainder)
    u32 W, Q, L, R, G;
    u32 classes_bytes;       // Q * sizeof(ClassT<RT, AT>)
    u32 smem_classes;        // 1: stage the class table in shared memory
    u32 smem_glist_cap;      // group-list entries staged in shared memory
    u32 smem_vorder;         // 1: stage vorder[] in shared memory
    u32 pack_enabled;
    // counts
    u32* total_local;        // [G] counts of this rank (zeroed here for the next tick)
    const u32* total_all;    // [G] counts summed over ranks (== total_local when not sharded)
    const u32* before;       // [G] counts of lower ranks, or nullptr
    // outputs
    GroupOut* gout;          // [G]
    u32* seg_cum;            // [SEG_CAP] inclusive end rank of the segment inside its group
    u32* seg_wv;             // [SEG_CAP] worker | variant << 16
    u64* free_after;         // [W][R]
    TickHeaderOut* hdr;
    uint2* glist;            // [G] scratch: non-empty groups (g, count) in processing order
    // peer-to-peer count exchange (sharded tick without a host collective): x_world == 0 => off
    const u32* x_counts;     // [x_world][HQS_MAX_GROUPS] count vectors written by the peers into MY exchange buffer
    const u32* x_flags;      // [x_world] tick sequence number each peer stores after its vector
    u32* x_all;              // [G] out: sum over ranks            (== total_all)
    u32* x_before;           // [G] out: sum over lower ranks      (== before)
    u32 x_world, x_rank, x_seq;
    // scan part
    u32* table;              // [P][G]
    u32 P;
    u32 scan_ctas;
    // pack part
    SolveSync* sync;
    PackScratch pk;
};

__device__ __forceinline__ u32 ld_acquire(const u32* p) {
    u32 v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(u32* p, u32 v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ u32 ld_acquire_sys(const u32* p) {
    u32 v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(u32* p, u32 v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Sharded tick, exchange step: block r stores this rank's per-group counts into peer r's exchange buffer (NVLink
// peer stores; r == own rank is a local copy) and then publishes the tick's sequence number with release
// semantics at system scope.  The peer's solver acquires the flag before it reads the vector.
struct XchgArgs {
    u32* peer[HQS_MAX_PEERS];     // base of every rank's exchange buffer (own included)
    u32 world, rank, seq, G;
};
__global__ void xchg_k(const u32* __restrict__ counts, XchgArgs x) {
    const u32 r = blockIdx.x;
    const u32 parity = x.seq & 1u;
    u32* dst = x.peer[r] + ((size_t)parity * HQS_MAX_PEERS + x.rank) * HQS_MAX_GROUPS;
    for (u32 g = threadIdx.x; g < x.G; g += blockDim.x) dst[g] = counts[g];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        u32* flags = x.peer[r] + (size_t)2 * HQS_MAX_PEERS * HQS_MAX_GROUPS + (size_t)parity * HQS_MAX_PEERS;
        st_release_sys(flags + x.rank, x.seq);
    }
}

template <int RT, typename AT>
__device__ __forceinline__ bool admissible(const VarT<RT, AT>& dv, u32 v, uint8_t blk, u64 rem_time) {
    return !((blk >> v) & 1) && (rem_time == HQS_TIME_INF || dv.min_time_ms <= rem_time);
}

// How many tasks of the variant fit into `fr` now, at most `cap` (< 2^32): min over the requested resources
// of floor(free / amount) (workerload.rs:121-145 without the 1024 cap).  `All`: feasible with >= 1 fraction
// (request.rs:34-36) but consumes the total (solver.rs:120-124), so at most one task and only on an untouched
// resource: bit r of `allok` says "total[r] != 0 and the parts of free/total the scaling dropped are equal".
// This sits on the solver's sequential critical path once per step, so it is STRAIGHT-LINE code: per resource
// (independent => ILP) a multiply-compare "does cap * amount fit" test and an fp32 quotient estimate with an
// exact integer fix-up, combined by selects.  Only a binding quotient of 2^20 or more (one worker taking over
// a million tasks of one group) falls back to an integer division.
template <int RT>
__device__ __forceinline__ u64 fit_count(const u64 (&fr)[RT], const u64 (&tot)[RT], u32 allok, const VarT<RT, u64>& dv,
                                         u64 cap) {
    u64 cnt = cap;
    bool big = false;
    const u32 used = dv.used_mask, allm = dv.all_mask;
#pragma unroll
    for (int r = 0; r < RT; ++r) {
        const bool on = (used >> r) & 1, all = (allm >> r) & 1;
        const u64 n = fr[r], d = dv.amount[r];
        const bool fits_cap = __umul64hi(d, cap) == 0 && d * cap <= n;
        // the 64-bit free amount is converted through its 32-bit halves (fp64 and 64-bit divisions cost hundreds of cycles)
        const float nf = __fmaf_rn(__uint2float_rn((u32)(n >> 32)), 4294967296.0f, __uint2float_rn((u32)n));
        const float qf = nf * dv.rcpf[r];
        u64 q = (u64)__float2uint_rz(fminf(qf, 1048576.0f));
        u64 p = q * d;
        q = p > n ? q - 1 : (n - p >= d ? q + 1 : q);
        p = q * d;
        q = p > n ? q - 1 : (n - p >= d ? q + 1 : q);
        const u64 q_all = (((allok >> r) & 1) && n == tot[r]) ? 1 : 0;
        const bool unconstrained = !on || (!all && (n == HQS_AMOUNT_MAX || fits_cap));
        big |= on && !all && !unconstrained && qf >= 1048576.0f;
        const u64 qr = all ? q_all : q;
        cnt = unconstrained ? cnt : (cnt < qr ? cnt : qr);
    }
    if (big) {                                  // rare: exact 64-bit divisions
        cnt = cap;
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            if (!((used >> r) & 1)) continue;
            u64 q;
            if ((allm >> r) & 1) q = (((allok >> r) & 1) && fr[r] == tot[r]) ? 1 : 0;
            else if (fr[r] != HQS_AMOUNT_MAX) q = fr[r] / dv.amount[r];
            else continue;
            cnt = cnt < q ? cnt : q;
        }
    }
    return cnt;
}

// Narrow amounts (< 2^31): one conversion, one multiply and a single +1 fix-up per resource.  The host stores
// rcpf = (1 / amount) * (1 - 2^-21): with every rounding counted the estimate is then never above the true
// quotient and, for quotients below 2^20, less than one below it, so floor(estimate) is q or q - 1.
template <int RT>
__device__ __forceinline__ u64 fit_count(const u32 (&fr)[RT], const u32 (&tot)[RT], u32 allok, const VarT<RT, u32>& dv,
                                         u64 cap64) {
    const u32 cap = (u32)cap64;
    u32 cnt = cap;
    bool big = false;
    const u32 used = dv.used_mask, allm = dv.all_mask;
    if (allm == 0) {
        // no `All` entry (the usual case): an unused resource has amount 0, which "fits cap" by itself
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            const u32 n = fr[r], d = dv.amount[r];
            const bool unconstrained = (u64)d * cap <= (u64)n || n == 0xFFFFFFFFu;
            const float qf = __uint2float_rn(n) * dv.rcpf[r];
            u32 q = __float2uint_rz(fminf(qf, 1048576.0f));
            q += (n - q * d >= d) ? 1u : 0u;
            big |= !unconstrained && qf >= 1048576.0f;
            cnt = unconstrained ? cnt : (cnt < q ? cnt : q);
        }
    } else {
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            const bool on = (used >> r) & 1, all = (allm >> r) & 1;
            const u32 n = fr[r], d = dv.amount[r];
            const bool fits_cap = (u64)d * cap <= (u64)n;
            const float qf = __uint2float_rn(n) * dv.rcpf[r];
            u32 q = __float2uint_rz(fminf(qf, 1048576.0f));
            q += (n - q * d >= d) ? 1u : 0u;
            const u32 q_all = (((allok >> r) & 1) && n == tot[r]) ? 1u : 0u;
            const bool unconstrained = !on || (!all && (n == 0xFFFFFFFFu || fits_cap));
            big |= on && !all && !unconstrained && qf >= 1048576.0f;
            const u32 qr = all ? q_all : q;
            cnt = unconstrained ? cnt : (cnt < qr ? cnt : qr);
        }
    }
    if (big) {                                  // rare: exact divisions
        cnt = cap;
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            if (!((used >> r) & 1)) continue;
            u32 q;
            if ((allm >> r) & 1) q = (((allok >> r) & 1) && fr[r] == tot[r]) ? 1u : 0u;
            else if (fr[r] != 0xFFFFFFFFu) q = fr[r] / dv.amount[r];
            else continue;
            cnt = cnt < q ? cnt : q;
        }
    }
    return cnt;
}

template <int RT, typename AT>
__device__ __forceinline__ void take_from(AT (&fr)[RT], const VarT<RT, AT>& dv, u64 k) {
#pragma unroll
    for (int r = 0; r < RT; ++r) {
        if (!((dv.used_mask >> r) & 1)) continue;
        if ((dv.all_mask >> r) & 1) fr[r] = 0;                               // workerload.rs:162
        else if (fr[r] != AmountMax<AT>::value) fr[r] -= (AT)k * dv.amount[r];
    }
}

// ---- pack: one warp fills one worker (specification: tests/greedy_model.py::_pack_level step b) ----
template <int RT>
__device__ void pack_body(const SolveArgs& a, unsigned char* smem_dyn) {
    // per-warp scratch in the (otherwise unused) dynamic shared memory of the pack CTAs
    double* s_dom = reinterpret_cast<double*>(smem_dyn) + (size_t)(threadIdx.x >> 5) * PACK_MAX_CAND;
    const ClassT<RT>* classes = reinterpret_cast<const ClassT<RT>*>(a.classes64);
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 n_pack_ctas = gridDim.x - 1;
    const u32 n_cand = __ldcg(a.pk.meta);
    // worker w is filled by CTA 1 + w % n_pack_ctas, warp w / n_pack_ctas: spreads the warps over the SMs
    for (u32 w = (blockIdx.x - 1) + warp * n_pack_ctas; w < a.W; w += n_pack_ctas * (blockDim.x >> 5)) {
        u64 fr[RT], tot[RT];
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            fr[r] = r < (int)a.R ? __ldcg(a.pk.fr + (size_t)w * a.R + r) : 0;
            tot[r] = r < (int)a.R ? a.total_rw[(size_t)w * a.R + r] : 0;
        }
        u32 allok = 0;
#pragma unroll
        for (int r = 0; r < RT; ++r) allok |= tot[r] != 0 ? (1u << r) : 0u;
        const u64 rem_time = a.rem_time[w];
        // exact u64 -> double through the 32-bit halves (one rounding, same value as a direct conversion)
        auto to_double = [](u64 x) -> double {
            return __dadd_rn(__dmul_rn(__uint2double_rn((u32)(x >> 32)), 4294967296.0), __uint2double_rn((u32)x));
        };
        // reciprocals once per worker / candidate: the per-iteration score is multiply-add only
        double inv_tot[RT];
#pragma unroll
        for (int r = 0; r < RT; ++r)
            inv_tot[r] = (tot[r] != 0 && tot[r] != HQS_AMOUNT_MAX) ? __ddiv_rn(1.0, to_double(tot[r])) : 0.0;
        // my two candidates
        u32 cls[2], var[2], gi[2], quota[2], taken[2], gs[2], ge[2];
        bool live[2];
        double inv_norm[2], dvec[2][RT];
        const VarT<RT>* dv[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const u32 ci = lane + 32 * j;
            live[j] = ci < n_cand;
            taken[j] = 0; quota[j] = 0; inv_norm[j] = 0.0; cls[j] = var[j] = gi[j] = 0; dv[j] = &classes[0].v[0];
            gs[j] = ge[j] = 0;
#pragma unroll
            for (int r = 0; r < RT; ++r) dvec[j][r] = 0.0;
            if (live[j]) {
                const u32 cd = __ldcg(a.pk.cand + ci);
                cls[j] = cd & 0xFFFFu; var[j] = (cd >> 16) & 0xFFu; gi[j] = cd >> 24;
                dv[j] = &classes[cls[j]].v[var[j]];
                gs[j] = ci - var[j];                                  // the variants of a group are consecutive candidates
                ge[j] = gs[j] + classes[cls[j]].n_variants;
                quota[j] = __ldcg(a.pk.quota + (size_t)w * PACK_MAX_CAND + gi[j]);
                const uint8_t blk = a.blocked ? a.blocked[(size_t)w * a.Q + cls[j]] : 0;
                live[j] = admissible(*dv[j], var[j], blk, rem_time);
                double s2 = 0.0;
#pragma unroll
                for (int r = 0; r < RT; ++r) {
                    if ((dv[j]->used_mask >> r) & 1) dvec[j][r] = __dmul_rn(to_double(dv[j]->amount[r]), inv_tot[r]);
                    s2 = __dadd_rn(s2, __dmul_rn(dvec[j][r], dvec[j][r]));
                }
                const double nrm = __dsqrt_rn(s2);
                inv_norm[j] = nrm > 0.0 ? __ddiv_rn(1.0, nrm) : 0.0;
            }
        }
        for (u32 it = 0; it < PACK_MAX_ITER; ++it) {
            double u[RT], inv_u[RT];
#pragma unroll
            for (int r = 0; r < RT; ++r) {
                u[r] = __dmul_rn(to_double(fr[r]), inv_tot[r]);
                inv_u[r] = __ddiv_rn(1.0, u[r]);                      // +inf where nothing is left
            }
            // a. per candidate: feasible? its dominant share of what the worker has left
            bool elig[2];
            double dom[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                elig[j] = live[j] && quota[j] != 0;
                dom[j] = 0.0;
                if (elig[j]) {
#pragma unroll
                    for (int r = 0; r < RT; ++r) {
                        if (((dv[j]->used_mask >> r) & 1) && fr[r] != HQS_AMOUNT_MAX && dv[j]->amount[r] > fr[r]) elig[j] = false;
                        if (dvec[j][r] > 0.0) {
                            const double x = __dmul_rn(dvec[j][r], inv_u[r]);
                            dom[j] = x > dom[j] ? x : dom[j];
                        }
                    }
                }
                if (lane + 32 * j < PACK_MAX_CAND) s_dom[lane + 32 * j] = elig[j] ? dom[j] : -1.0;   // -1: not eligible
            }
            __syncwarp();
            // b. per group the eligible variant with the smallest share (ties: lower index) stays in the race
            double best_s = 0.0;
            u32 best_ci = ~0u;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (!elig[j]) continue;
                const u32 ci = lane + 32 * j;
                bool win = true;
                for (u32 k = gs[j]; k < ge[j]; ++k) {
                    const double o = s_dom[k];
                    if (k != ci && o >= 0.0 && (o < dom[j] || (o == dom[j] && k < ci))) win = false;
                }
                if (!win) continue;
                double dot = 0.0;
#pragma unroll
                for (int r = 0; r < RT; ++r) dot = __dadd_rn(dot, __dmul_rn(dvec[j][r], u[r]));
                const double sc = __dmul_rn(dot, inv_norm[j]);
                if (best_ci == ~0u || sc > best_s) { best_s = sc; best_ci = ci; }   // j = 0 first: lower index wins ties
            }
            __syncwarp();
            // c. warp argmax: larger score, ties to the lower candidate index
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
                const double os = __shfl_xor_sync(0xffffffffu, best_s, d);
                const u32 oc = __shfl_xor_sync(0xffffffffu, best_ci, d);
                if (oc != ~0u && (best_ci == ~0u || os > best_s || (os == best_s && oc < best_ci))) { best_s = os; best_ci = oc; }
            }
            if (best_ci == ~0u) break;
            const u32 owner = best_ci & 31, oj = best_ci >> 5;
            u32 k = 0, ggi = 0;
            const VarT<RT>* mydv = oj ? dv[1] : dv[0];
            if (lane == owner) {
                const u32 q = oj ? quota[1] : quota[0];
                const u64 f = fit_count<RT>(fr, tot, allok, *mydv, q);
                const u32 chunk = q / PACK_CHUNK_DIV > 1 ? q / PACK_CHUNK_DIV : 1;
                k = (u32)(f < chunk ? f : chunk);
                if (oj) taken[1] += k; else taken[0] += k;
                ggi = oj ? gi[1] : gi[0];
            }
            k = __shfl_sync(0xffffffffu, k, owner);
            ggi = __shfl_sync(0xffffffffu, ggi, owner);
            // every lane applies the owner's amounts to its copy of the free vector
            const VarT<RT>* odv = (const VarT<RT>*)__shfl_sync(0xffffffffu, (unsigned long long)mydv, owner);
            take_from<RT, u64>(fr, *odv, k);
#pragma unroll
            for (int j = 0; j < 2; ++j)
                if (gi[j] == ggi && (lane + 32 * j) < n_cand) quota[j] = quota[j] >= k ? quota[j] - k : 0;
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)
            if (lane + 32 * j < n_cand) a.pk.taken[(size_t)w * PACK_MAX_CAND + lane + 32 * j] = taken[j];
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < RT; ++r)
                if (r < (int)a.R) a.pk.fr[(size_t)w * a.R + r] = fr[r];
        }
    }
}

// ---- CTA 0 ---------------------------------------------------------------------------------------
struct ScanOut {
    u32 take, exc_cnt, n_takers;
    u64 tot_cnt;
};

// Block-wide "hand out `remaining` units in worker order": thread (worker) w offers cnt, receives
// take = clamp(remaining - sum_{w' < w} cnt_{w'}, 0, cnt).  One barrier (double-buffered exchange): warp
// inclusive scan by shuffles, then every thread adds up the (few) warp totals below it serially — 8 loads
// and adds beat a second 5-step shuffle scan on this latency-bound path.  The rank of a taker among the
// workers that offer anything comes from a ballot, not from the scan.
template <int MAXW>
__device__ __forceinline__ ScanOut scan_take(u64 cnt, u32 remaining, u64* s_x, u32& parity, u32& seg_rank) {
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64* buf = s_x + 32 * (parity & 1);
    parity++;
    u64 inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u64 y = __shfl_up_sync(0xffffffffu, inc, d);
        if ((int)lane >= d) inc += y;
    }
    const u32 hasb = __ballot_sync(0xffffffffu, cnt != 0);
    if (lane == 31) buf[warp] = inc | ((u64)__popc(hasb) << 42);       // low 42 bits count, high bits offerers
    __syncthreads();
    u64 below = 0, all = 0;
#pragma unroll
    for (int w2 = 0; w2 < MAXW; ++w2) {          // slots of warps that do not exist stay zero (cleared at start)
        const u64 v = buf[w2];
        all += v;
        below += (u32)w2 < warp ? v : 0ull;
    }
    const u64 mask = (1ull << 42) - 1;
    const u64 exc = (below & mask) + inc - cnt;
    ScanOut o;
    o.exc_cnt = (u32)(exc < remaining ? exc : remaining);
    o.take = 0;
    if (cnt && exc < remaining) {
        const u64 room = remaining - exc;
        o.take = (u32)(cnt < room ? cnt : room);
    }
    seg_rank = (u32)(below >> 42) + __popc(hasb & ((1u << lane) - 1));
    o.tot_cnt = all & mask;
    o.n_takers = 0;      // the caller counts the takers (a second barrier that also fences s_x reuse)
    return o;
}

template <int RT, int MAXT, bool SMALL, typename AT>
__device__ void solve_body(const SolveArgs& a, unsigned char* smem_dyn) {
    constexpr bool NARROW = sizeof(AT) == 4;
    constexpr AT AMAX = AmountMax<AT>::value;
    using Var = VarT<RT, AT>;
    using Cls = ClassT<RT, AT>;
    __shared__ u64 s_x[64], s_f[64];
    __shared__ u64 s_red[2 * 32 * (2 * RT + 1)];
    __shared__ u64 s_totmax[RT];
    __shared__ u32 s_a[40], s_b[40];
    __shared__ u32 s_nlist;
    __shared__ u64 s_qsum[PACK_MAX_CAND * (MAXT / 32)];     // packed level: warp sums of the per-worker fit counts
    const u32 tid = threadIdx.x;
    const u32 lane = tid & 31, warp = tid >> 5;
    const u32 nwarps = blockDim.x >> 5;
    const bool has_worker = tid < a.W;
    u32 parity = 0;
    if (tid < 64) s_x[tid] = 0;          // scan_take sums a fixed number of warp slots
    if (tid < 40) s_b[tid] = 0;

    // ---- class table and variant order: in shared memory when they fit (SMALL: the pointers are then
    //      provably shared, so the sequential critical path uses LDS, not generic loads), else global
    const Cls* classes;
    const uint8_t* vorder;
    unsigned char* sp = smem_dyn;
    if constexpr (SMALL) {
        const uint4* src = reinterpret_cast<const uint4*>(a.classes);
        uint4* dst = reinterpret_cast<uint4*>(smem_dyn);
        for (u32 i = tid; i < a.classes_bytes / 16; i += blockDim.x) dst[i] = src[i];
        classes = reinterpret_cast<const Cls*>(smem_dyn);
        sp += (a.classes_bytes + 15u) & ~15u;
        uint8_t* sv = sp;
        for (u32 i = tid; i < a.Q * HQS_MAX_VARIANTS; i += blockDim.x) sv[i] = a.vorder[i];
        vorder = sv;
        sp += (a.Q * HQS_MAX_VARIANTS + 15u) & ~15u;
    } else {
        classes = reinterpret_cast<const Cls*>(a.classes);
        vorder = a.vorder;
    }
    uint2* s_glist = reinterpret_cast<uint2*>(sp);
    u32* s_gcl = reinterpret_cast<u32*>(sp + (size_t)a.smem_glist_cap * sizeof(uint2));   // [gl_cap] class | level << 16
    u32* s_conf = reinterpret_cast<u32*>(sp + (size_t)a.smem_glist_cap * (sizeof(uint2) + sizeof(u32)));   // bit e: group e confirmed placeable this round
    const u32 n_conf_words = (a.smem_glist_cap + 31) / 32;
    uint8_t* s_alive = reinterpret_cast<uint8_t*>(s_conf + n_conf_words);                 // [gl_cap] group still placeable
    // Outputs of the sequential loop are buffered in shared memory and written out in bulk: a global store
    // in front of a barrier costs an L2 round trip per step (bar.sync waits for the store to be visible).
    unsigned char* sp2 = sp + (((size_t)a.smem_glist_cap * (sizeof(uint2) + sizeof(u32) + 1) + (size_t)n_conf_words * 4 + 15) & ~size_t(15));
    GroupOut* s_gout = reinterpret_cast<GroupOut*>(sp2);                                   // [gl_cap], by entry
    u32* s_segc = reinterpret_cast<u32*>(sp2 + (size_t)a.smem_glist_cap * sizeof(GroupOut)); // [SEG_SMEM]
    u32* s_segw = s_segc + SEG_SMEM;
    u32 seg_flushed = 0;                                                                    // uniform
    const u32 gl_cap = a.smem_glist_cap;
#define GL(e) ((e) < gl_cap ? s_glist[(e)] : a.glist[(e)])
#define GC(e) ((e) < gl_cap ? (s_gcl[(e)] & 0xFFFFu) : (a.glist[(e)].x % a.Q))      /* class of entry e */
#define GLV(e) ((e) < gl_cap ? (s_gcl[(e)] >> 16) : (a.glist[(e)].x / a.Q))        /* level of entry e */

    bool x_timeout = false;
    if (a.x_world) {
        __syncthreads();
        // sharded tick: wait until every rank's count vector of THIS tick has landed in my exchange buffer, then
        // materialise sum-over-ranks and sum-over-lower-ranks (the two vectors the host all-gather used to provide)
        if (tid < a.x_world) {
            const long long t0 = clock64();
            while (ld_acquire_sys(a.x_flags + tid) != a.x_seq) {
                if (clock64() - t0 > SPIN_TIMEOUT_CYCLES) { s_b[1] = 1; break; }
                __nanosleep(32);
            }
        }
        __syncthreads();
        for (u32 g = tid; g < a.G; g += blockDim.x) {
            u32 all = 0, bef = 0;
            for (u32 r = 0; r < a.x_world; ++r) {
                const u32 v = __ldcg(a.x_counts + (size_t)r * HQS_MAX_GROUPS + g);     // peers wrote it: bypass L1
                all += v;
                bef += r < a.x_rank ? v : 0u;
            }
            a.x_all[g] = all;
            a.x_before[g] = bef;
        }
        __syncthreads();
        x_timeout = s_b[1] == 1;
    }
    const long long t_start = clock64();
    long long t_sat = 0, t_groups = 0;
    // worker state in registers.  Narrow path: fr/tot hold floor(amount / gscale[r]), `rem` what the division
    // dropped (exact amount = fr * gscale + rem; requests are multiples of gscale, so rem only changes when an
    // `All` request empties the resource).
    AT fr[RT], tot[RT];
    u64 rem[NARROW ? RT : 1];
    u32 allok = 0;      // bit r: total != 0 and free/total agree in the dropped part (-> `All` needs fr == tot)
#pragma unroll
    for (int r = 0; r < RT; ++r) {
        const u64 n = (has_worker && r < (int)a.R) ? a.free_rw[(size_t)tid * a.R + r] : 0;
        const u64 t = (has_worker && r < (int)a.R) ? a.total_rw[(size_t)tid * a.R + r] : 0;
        if constexpr (NARROW) {
            const u64 g = a.gscale[r];
            const u64 nq = g == 1 ? n : n / g, tq = g == 1 ? t : t / g;
            fr[r] = n == HQS_AMOUNT_MAX ? AMAX : (AT)nq;
            tot[r] = t == HQS_AMOUNT_MAX ? AMAX : (AT)tq;
            rem[r] = n == HQS_AMOUNT_MAX ? 0 : n - nq * g;
            const u64 trem = t == HQS_AMOUNT_MAX ? 0 : t - tq * g;
            allok |= (t != 0 && rem[r] == trem) ? (1u << r) : 0u;
        } else {
            fr[r] = n; tot[r] = t;
            allok |= t != 0 ? (1u << r) : 0u;
        }
    }
    auto exact_free = [&](int r) -> u64 {
        if constexpr (NARROW) return fr[r] == AMAX ? HQS_AMOUNT_MAX : (u64)fr[r] * a.gscale[r] + rem[r];
        else return fr[r];
    };
    auto exact_amount = [&](const Var& dv, int r) -> u64 {
        if constexpr (NARROW) return (u64)dv.amount[r] * a.gscale[r];
        else return dv.amount[r];
    };
    const u64 rem_time = has_worker ? a.rem_time[tid] : 0;
    // per-resource maximum of the worker totals (a class no worker is big enough for is not demand)
    {
        u64 m[RT];
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            m[r] = (u64)tot[r];
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) { const u64 y = __shfl_xor_sync(0xffffffffu, m[r], d); m[r] = y > m[r] ? y : m[r]; }
        }
        if (lane == 0)
#pragma unroll
            for (int r = 0; r < RT; ++r) s_red[warp * RT + r] = m[r];
        __syncthreads();
        if (tid < RT) {
            u64 best = 0;
            for (u32 w2 = 0; w2 < nwarps; ++w2) best = s_red[w2 * RT + tid] > best ? s_red[w2 * RT + tid] : best;
            s_totmax[tid] = best;
        }
        __syncthreads();
    }

    // ---- compact the non-empty groups, in processing order: level asc (= priority desc), then the
    //      tick's class order
    if (tid == 0) s_nlist = 0;
    __syncthreads();
    const u32 n_pos = a.L * a.Q;
    for (u32 base = 0; base < n_pos; base += blockDim.x) {
        const u32 pos = base + tid;
        u32 g = 0, n = 0;
        if (pos < n_pos) {
            const u32 lvl = pos / a.Q, j = pos - lvl * a.Q;
            g = lvl * a.Q + a.order[j];
            n = a.total_all[g];
        }
        const u32 bal = __ballot_sync(0xffffffffu, n != 0);
        if (lane == 0) s_a[warp] = __popc(bal);
        __syncthreads();
        u32 off = s_nlist;
        for (u32 w2 = 0; w2 < warp; ++w2) off += s_a[w2];
        if (n) {
            const u32 slot = off + __popc(bal & ((1u << lane) - 1));
            a.glist[slot] = make_uint2(g, n);
            if (slot < gl_cap) { s_glist[slot] = make_uint2(g, n); s_gcl[slot] = (g % a.Q) | ((g / a.Q) << 16); }
        }
        __syncthreads();
        if (tid == 0) {
            u32 t = 0;
            for (u32 w2 = 0; w2 < nwarps; ++w2) t += s_a[w2];
            s_nlist += t;
        }
        __syncthreads();
    }
    const u32 n_list = s_nlist;
    for (u32 e = tid; e < gl_cap; e += blockDim.x) {
        s_alive[e] = 1;
        if (e < n_conf_words) s_conf[e] = 0;
        GroupOut z; z.k = 0; z.out_off = 0; z.seg_lo = 0; z.seg_n = 0;
        s_gout[e] = z;
    }
#define FLUSH_SEGMENTS_IF_FULL()                                                                      \
    if (seg_base - seg_flushed + blockDim.x > SEG_SMEM) {                                             \
        __syncthreads();                                                                              \
        for (u32 i_ = tid; i_ < seg_base - seg_flushed; i_ += blockDim.x) {                          \
            a.seg_cum[seg_flushed + i_] = s_segc[i_];                                                 \
            a.seg_wv[seg_flushed + i_] = s_segw[i_];                                                  \
        }                                                                                             \
        __syncthreads();                                                                              \
        seg_flushed = seg_base;                                                                       \
    }
    // groups without ready tasks keep k = 0 (emit_k's chunk filter reads k of every group)
    for (u32 g = tid; g < a.G; g += blockDim.x) a.gout[g].k = 0;
    __syncthreads();
    const long long t_compact = clock64();

    u32 seg_base = 0;    // uniform across the CTA
    u32 out_base = 0;    // uniform: local output offset
    bool seg_overflow = false, sync_timeout = false;
    bool packed = a.pack_enabled == 0;
    bool signalled = false;

    u32 li = 0;
    while (li < n_list) {
        // ---- one priority level: entries [li, lj)
        const u32 lvl = GLV(li);
        u32 lj = li + 1;
        while (lj < n_list && GLV(lj) == lvl) ++lj;
        const u32 ng = lj - li;
        bool level_packed = false;
        const long long t_l0 = clock64();

        if (!packed && ng <= PACK_MAX_CAND) {
            // ---- is this level saturated?  demand (first variant of the tick's order) vs free, exact
            //      saturating u64.  Thread t < ng owns group li + t; one exchange reduces, per resource, the
            //      free capacity over workers and the demand over groups, plus the candidate count.
            constexpr int NV = 2 * RT + 1;
            u64 val[NV];
#pragma unroll
            for (int r = 0; r < RT; ++r) { val[r] = has_worker ? exact_free(r) : 0; val[RT + r] = 0; }
            val[2 * RT] = 0;
            if (tid < ng) {
                const uint2 ge = GL(li + tid);
                const u32 c = GC(li + tid);
                const u32 nvv = classes[c].n_variants;
                u64 flag = 0;
                for (u32 v = 0; v < nvv; ++v) flag |= classes[c].v[v].all_mask ? (1ull << 32) : 0ull;
                val[2 * RT] = nvv | flag;                                  // low: candidates, bit 32+: has `All`
                const Var& dv = classes[c].v[vorder[c * HQS_MAX_VARIANTS]];
                bool servable = true;
#pragma unroll
                for (int r = 0; r < RT; ++r) servable &= (u64)dv.amount[r] <= s_totmax[r];
#pragma unroll
                for (int r = 0; r < RT; ++r) {
                    const u64 amt = exact_amount(dv, r);
                    const u64 hi = __umul64hi(amt, (u64)ge.y);
                    val[RT + r] = !servable ? 0 : (hi ? HQS_AMOUNT_MAX : amt * (u64)ge.y);
                }
            }
#pragma unroll
            for (int i = 0; i < NV; ++i) {
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) {
                    const u64 y = __shfl_xor_sync(0xffffffffu, val[i], d);
                    const u64 sum = val[i] + y;
                    val[i] = sum < y ? HQS_AMOUNT_MAX : sum;              // saturating (MAX absorbs)
                }
            }
            u64* red = s_red + (size_t)(parity & 1) * 32 * NV;
            parity++;
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < NV; ++i) red[warp * NV + i] = val[i];
            }
            __syncthreads();
            u64 tot_v = 0;                                                 // lane i < NV sums column i over warps
            if (lane < NV)
                for (u32 w2 = 0; w2 < nwarps; ++w2) {
                    const u64 y = red[w2 * NV + lane];
                    const u64 sum = tot_v + y;
                    tot_v = sum < y ? HQS_AMOUNT_MAX : sum;
                }
            const u64 meta = __shfl_sync(0xffffffffu, tot_v, 2 * RT);
            const u32 n_cand = (u32)(meta & 0xFFFFFFFFu);
            const bool has_all = (meta >> 32) != 0;
            if (n_cand <= PACK_MAX_CAND && !has_all) {
                // phi = the fraction of the level's demand the pool can serve, when two or more resources are
                // over-subscribed (the classes then complement each other and each gets the same fraction of
                // its demand this tick); with a single scarce resource any split drains at the same rate
                u32 n_sat = 0;
                double phi = 1.0;
                for (u32 r = 0; r < a.R; ++r) {
                    const u64 C = __shfl_sync(0xffffffffu, tot_v, r);
                    const u64 D = __shfl_sync(0xffffffffu, tot_v, RT + r);
                    if (C != HQS_AMOUNT_MAX && D > C) n_sat++;
                }
                if (n_sat >= 2)
                    for (u32 r = 0; r < a.R; ++r) {
                        const u64 C = __shfl_sync(0xffffffffu, tot_v, r);
                        const u64 D = __shfl_sync(0xffffffffu, tot_v, RT + r);
                        if (C != HQS_AMOUNT_MAX && D > 0) {
                            const double x = __ddiv_rn(__ull2double_rn(C), __ull2double_rn(D));
                            phi = x < phi ? x : phi;
                        }
                    }
                const bool saturated = n_sat != 0;
                if (saturated) {
                    // ---- a. quotas: share of each class proportional to how many fit on the worker alone.
                    //      Pass 1: every worker's own count per group (stashed in its quota slot) and the warp
                    //      sums; ONE barrier; pass 2: pool totals and the quotas.
                    constexpr u32 NW = MAXT / 32;
                    for (u32 e = li; e < lj; ++e) {
                        const uint2 ge = GL(e);
                        const u32 c = GC(e), n = ge.y;
                        const uint8_t blk = (has_worker && a.blocked) ? a.blocked[(size_t)tid * a.Q + c] : 0;
                        u64 cn = 0;
                        if (has_worker) {
                            for (u32 v = 0; v < classes[c].n_variants; ++v) {
                                const Var& dv = classes[c].v[v];
                                if (!admissible(dv, v, blk, rem_time)) continue;
                                const u64 f = fit_count<RT>(fr, tot, allok, dv, n);
                                cn = f > cn ? f : cn;
                            }
                            a.pk.quota[(size_t)tid * PACK_MAX_CAND + (e - li)] = (u32)cn;       // cn <= n < 2^32
                        }
                        u64 x = cn;
#pragma unroll
                        for (int d = 16; d >= 1; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
                        if (lane == 0) s_qsum[(e - li) * NW + warp] = x;
                    }
                    __syncthreads();
                    for (u32 e = li; e < lj; ++e) {
                        const u32 n = GL(e).y;
                        u64 T = 0;
                        for (u32 w2 = 0; w2 < nwarps; ++w2) T += s_qsum[(e - li) * NW + w2];
                        if (has_worker) {
                            const u64 cn = a.pk.quota[(size_t)tid * PACK_MAX_CAND + (e - li)];   // this thread's own store
                            const u64 q = T ? ((u64)n * cn + T - 1) / T : 0;
                            const u64 q_phi = __double2ull_ru(__dmul_rn(__ull2double_rn(q), phi));     // ceil(q * phi)
                            a.pk.quota[(size_t)tid * PACK_MAX_CAND + (e - li)] = (u32)q_phi;
                        }
                    }
                    // ---- b. publish the worker state and the candidate list, release the pack warps
                    if (has_worker) {
#pragma unroll
                        for (int r = 0; r < RT; ++r)
                            if (r < (int)a.R) a.pk.fr[(size_t)tid * a.R + r] = exact_free(r);
                    }
                    if (tid == 0) {
                        u32 ci = 0;
                        for (u32 e = li; e < lj; ++e) {
                            const u32 c = GC(e);
                            for (u32 v = 0; v < classes[c].n_variants; ++v) a.pk.cand[ci++] = c | (v << 16) | ((e - li) << 24);
                        }
                        a.pk.meta[0] = ci;
                        a.pk.meta[1] = ng;
                    }
                    __threadfence();
                    __syncthreads();
                    if (tid == 0) {
                        u32 timed_out = 0;
                        st_release(&a.sync->phase, PHASE_PACK);
                        const long long t0 = clock64();
                        while (ld_acquire(&a.sync->done) < gridDim.x - 1) {
                            if (clock64() - t0 > SPIN_TIMEOUT_CYCLES) { timed_out = 1; break; }
                            __nanosleep(64);
                        }
                        s_b[0] = timed_out;
                    }
                    __syncthreads();
                    sync_timeout = s_b[0] == 1;
                    signalled = true;
                    if (has_worker) {
#pragma unroll
                        for (int r = 0; r < RT; ++r)
                            if (r < (int)a.R) {
                                const u64 x = __ldcg(a.pk.fr + (size_t)tid * a.R + r);
                                if constexpr (NARROW) {
                                    const u64 g = a.gscale[r];
                                    fr[r] = x == HQS_AMOUNT_MAX ? AMAX : (AT)(g == 1 ? x - rem[r] : (x - rem[r]) / g);
                                } else {
                                    fr[r] = x;
                                }
                            }
                    }
                    packed = true;
                    level_packed = !sync_timeout;
                }
            }
        }

        const long long t_l1 = clock64();
        t_sat += t_l1 - t_l0;
        // ---- the groups of the level, in order: cap what pack took, then first-fit the rest
        u32 cand_base = 0;
        bool level_unsatisfied = false;
        for (u32 e = li; e < lj; ++e) {
            if (e < gl_cap && !s_alive[e]) {                  // no worker can take a single task of it (uniform)
                if (level_packed) cand_base += classes[GC(e)].n_variants;
                continue;
            }
            const uint2 ge = GL(e);
            const u32 g = ge.x, n_all = ge.y;
            const u32 c = GC(e);
            const u32 nv = classes[c].n_variants;
            u32 remaining = n_all;
            const u32 seg_lo = seg_base;
            const uint8_t blk = (has_worker && a.blocked) ? a.blocked[(size_t)tid * a.Q + c] : 0;
            if (level_packed) {
                for (u32 v = 0; v < nv; ++v) {
                    FLUSH_SEGMENTS_IF_FULL();
                    const u64 cnt = has_worker ? __ldcg(a.pk.taken + (size_t)tid * PACK_MAX_CAND + cand_base + v) : 0;
                    u32 seg_rank;
                    ScanOut o = scan_take<MAXT / 32>(cnt, remaining, s_x, parity, seg_rank);
                    const Var& dv = classes[c].v[v];
                    if (o.take) {
                        const u32 si = seg_base + seg_rank;
                        if (si < SEG_CAP) {
                            s_segc[si - seg_flushed] = (n_all - remaining) + o.exc_cnt + o.take;
                            s_segw[si - seg_flushed] = tid | (v << 16);
                        }
                    }
                    if (cnt > o.take) {
                        const u64 ex = cnt - o.take;
#pragma unroll
                        for (int r = 0; r < RT; ++r)
                            if (((dv.used_mask >> r) & 1) && fr[r] != AMAX) fr[r] += (AT)ex * dv.amount[r];
                    }
                    const u32 n_takers = (u32)__syncthreads_count(o.take != 0);
                    remaining -= (u32)(o.tot_cnt < remaining ? o.tot_cnt : remaining);
                    seg_base += n_takers;
                    if (seg_base > SEG_CAP) { seg_overflow = true; seg_base = SEG_CAP; }
                }
                cand_base += nv;
            }
            u32 tried = 0;                                            // per worker: variants of this class already offered
            for (u32 vi = 0; vi < nv && remaining > 0; ++vi) {
                // Each worker offers the untried variant that costs the smallest share of what it has left:
                // min over variants of max_r f32(amount_r) * (1 / f32(free_r)), `All` = +inf, ties to the lower
                // variant id (specification: tests/greedy_model.py::_Tick.next_variant).
                u32 v = 0;
                if (nv > 1) {
                    float inv[RT];
#pragma unroll
                    for (int r = 0; r < RT; ++r) inv[r] = __fdiv_rn(1.0f, __double2float_rn(__ull2double_rn(exact_free(r))));
                    float best_d = 0.0f;
                    int best_v = -1;
                    for (u32 vv = 0; vv < nv; ++vv) {
                        if ((tried >> vv) & 1) continue;
                        const Var& cv = classes[c].v[vv];
                        float dom = 0.0f;
                        if (cv.all_mask) dom = __int_as_float(0x7f800000);
                        else {
#pragma unroll
                            for (int r = 0; r < RT; ++r) {
                                if (!((cv.used_mask >> r) & 1) || fr[r] == AMAX) continue;
                                const float x = __fmul_rn(cv.rcpf[RT + r], inv[r]);
                                dom = x > dom ? x : dom;
                            }
                        }
                        if (best_v < 0 || dom < best_d) { best_v = (int)vv; best_d = dom; }
                    }
                    v = (u32)best_v;
                    tried |= 1u << v;
                }
                const Var& dv = classes[c].v[v];
                FLUSH_SEGMENTS_IF_FULL();
                // exact count once (reciprocal division, capped at `remaining`): can1 = cnt > 0, and the
                // first worker takes everything iff its cnt == remaining
                u64 cnt = 0;
                if (has_worker && admissible(dv, v, blk, rem_time)) cnt = fit_count<RT>(fr, tot, allok, dv, remaining);
                const bool can1 = cnt != 0, can_all = cnt >= remaining;
                u32 take = 0, exc_cnt = 0, seg_rank = 0, n_takers = 0, handed = 0;
                {
                    u64* fb = s_f + 32 * (parity & 1);
                    parity++;
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
                        if (ef & 1ull) {                       // the first worker that can take anything takes it all
                            take = (warp == wf && lane == first && can1) ? remaining : 0;
                            n_takers = 1; handed = remaining;
                        } else {
                            ScanOut o = scan_take<MAXT / 32>(cnt, remaining, s_x, parity, seg_rank);
                            take = o.take; exc_cnt = o.exc_cnt;
                            n_takers = (u32)__syncthreads_count(o.take != 0);
                            handed = (u32)(o.tot_cnt < remaining ? o.tot_cnt : remaining);
                        }
                    }
                }
                if (take) {
                    const u32 si = seg_base + seg_rank;
                    if (si < SEG_CAP) {
                        s_segc[si - seg_flushed] = (n_all - remaining) + exc_cnt + take;
                        s_segw[si - seg_flushed] = tid | (v << 16);
                    }
                    take_from<RT, AT>(fr, dv, take);
                    if constexpr (NARROW) {
                        // `All` consumed the whole resource: the exact free amount is 0, remainder included
                        const u32 z = dv.all_mask & dv.used_mask;
                        if (z) {
#pragma unroll
                            for (int r = 0; r < RT; ++r)
                                if ((z >> r) & 1) rem[r] = 0;
                            allok &= ~z;
                        }
                    }
                }
                remaining -= handed;
                seg_base += n_takers;
                if (seg_base > SEG_CAP) { seg_overflow = true; seg_base = SEG_CAP; }
            }
            const u32 k = n_all - remaining;
            level_unsatisfied |= remaining != 0;
            // local share of the k assigned tasks (sharded mode: ranks are ordered by handle range)
            u32 k_loc = k;
            if (a.before) {
                const u32 bef = a.before[g], loc = a.total_local[g];
                k_loc = k > bef ? k - bef : 0;
                k_loc = k_loc < loc ? k_loc : loc;
            }
            if (tid == 0) {
                GroupOut go;
                go.k = k; go.out_off = out_base; go.seg_lo = seg_lo; go.seg_n = seg_base - seg_lo;
                if (e < gl_cap) s_gout[e] = go; else a.gout[g] = go;
            }
            out_base += k_loc;
        }
        // ---- the level left tasks behind: the pool is filling up.  Free amounts only shrink from here on
        //      (pack's hand-backs are over), so a later group that no worker can take one task of NOW can be
        //      dropped for the rest of the tick: test all of them at once, one thread per worker.
        if (level_unsatisfied && lj < n_list && n_list <= gl_cap) {
            for (u32 e = lj; e < n_list; ++e) {
                if (!s_alive[e]) continue;                                     // uniform
                const u32 c = GC(e);
                const uint8_t blk = (has_worker && a.blocked) ? a.blocked[(size_t)tid * a.Q + c] : 0;
                bool can = false;
                if (has_worker)
                    for (u32 v = 0; v < classes[c].n_variants && !can; ++v) {
                        const Var& dv = classes[c].v[v];
                        if (!admissible(dv, v, blk, rem_time)) continue;
                        bool ok = true;
#pragma unroll
                        for (int r = 0; r < RT; ++r) {
                            if (!((dv.used_mask >> r) & 1)) continue;
                            if ((dv.all_mask >> r) & 1) ok &= ((allok >> r) & 1) && fr[r] == tot[r];
                            else if (fr[r] != AMAX) ok &= dv.amount[r] <= fr[r];
                        }
                        can = ok;
                    }
                const u32 anyc = __ballot_sync(0xffffffffu, can);
                if (lane == 0 && anyc) atomicOr(&s_conf[e >> 5], 1u << (e & 31));
            }
            __syncthreads();
            for (u32 e = lj + tid; e < n_list; e += blockDim.x) s_alive[e] = s_alive[e] && ((s_conf[e >> 5] >> (e & 31)) & 1);
            __syncthreads();
            for (u32 i = tid; i < n_conf_words; i += blockDim.x) s_conf[i] = 0;      // next use is behind later barriers
        }
        t_groups += clock64() - t_l1;
        li = lj;
    }
    const long long t_loop = clock64();

    // ---- flush the buffered segments and per-group records
    __syncthreads();
    for (u32 i = tid; i < seg_base - seg_flushed; i += blockDim.x) {
        a.seg_cum[seg_flushed + i] = s_segc[i];
        a.seg_wv[seg_flushed + i] = s_segw[i];
    }
    for (u32 e = tid; e < n_list && e < gl_cap; e += blockDim.x) a.gout[s_glist[e].x] = s_gout[e];
    // ---- epilogue: let the other CTAs go, header, free vectors after the tick, reset the counters
    if (tid == 0 && !signalled) st_release(&a.sync->phase, PHASE_EXIT);
    if (has_worker) {
#pragma unroll
        for (int r = 0; r < RT; ++r)
            if (r < (int)a.R) a.free_after[(size_t)tid * a.R + r] = exact_free(r);
    }
    if (tid == 0) {
        a.hdr->n_assigned = out_base;
        a.hdr->n_groups = n_list;
        a.hdr->n_segments = seg_base;
        a.hdr->error = (sync_timeout || x_timeout) ? 2u : (seg_overflow ? 1u : 0u);
        a.hdr->dbg[0] = t_compact - t_start; a.hdr->dbg[1] = t_sat; a.hdr->dbg[2] = t_groups;
        a.hdr->dbg[3] = t_loop - t_start; a.hdr->dbg[4] = n_list;
    }
    __syncthreads();
    for (u32 g = tid; g < a.G; g += blockDim.x) a.total_local[g] = 0;
#undef GL
#undef GC
#undef GLV
#undef FLUSH_SEGMENTS_IF_FULL
}

template <int RT, int MAXT, bool SMALL, typename AT>
__global__ void __launch_bounds__(MAXT) solve_k(SolveArgs a) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    if (blockIdx.x == 0) {
        solve_body<RT, MAXT, SMALL, AT>(a, smem_dyn);
        return;
    }
    // ---- exclusive scan over chunks: one warp per group column, 32 chunk rows per step
    if (blockIdx.x <= a.scan_ctas) {
        const u32 lane = threadIdx.x & 31;
        const u32 nw = blockDim.x >> 5;
        for (u32 g = (blockIdx.x - 1) * nw + (threadIdx.x >> 5); g < a.G; g += a.scan_ctas * nw) {
            u32 carry = 0;
            for (u32 b0 = 0; b0 < a.P; b0 += 32) {
                const u32 b = b0 + lane;
                const u32 v = b < a.P ? a.table[(size_t)b * a.G + g] : 0;
                u32 inc = v;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const u32 y = __shfl_up_sync(0xffffffffu, inc, d);
                    if ((int)lane >= d) inc += y;
                }
                if (b < a.P) a.table[(size_t)b * a.G + g] = carry + inc - v;
                carry += __shfl_sync(0xffffffffu, inc, 31);
            }
        }
    }
    if (!a.pack_enabled) return;            // nothing to stand by for
    // ---- wait for CTA 0's decision
    __shared__ u32 s_cmd;
    if (threadIdx.x == 0) {
        u32 cmd;
        const long long t0 = clock64();
        while ((cmd = ld_acquire(&a.sync->phase)) == PHASE_WAIT) {
            if (clock64() - t0 > SPIN_TIMEOUT_CYCLES) { cmd = PHASE_EXIT; break; }
            __nanosleep(128);
        }
        s_cmd = cmd;
    }
    __syncthreads();
    if (s_cmd == PHASE_PACK) {
        pack_body<RT>(a, smem_dyn);
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(&a.sync->done, 1u);
    }
}

