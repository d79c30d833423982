// hqsched.cu — B200 (sm_100a) task->worker assignment solver behind the C ABI of include/hqsched.h.
//
// Replaces, for the single-node hot path of HyperQueue's tako scheduler tick (v0.26.0):
//   create_task_batches        crates/tako/src/internal/scheduler/batches.rs:42-181   (priority histogram)
//   run_scheduling_solver      crates/tako/src/internal/scheduler/solver.rs:16-461    (who gets how many)
//   create_task_mapping        crates/tako/src/internal/scheduler/mapping.rs:23-154   (which task goes where)
//   TaskQueues                 crates/tako/src/internal/scheduler/taskqueue.rs        (ready set, device resident)
//   task_finished readiness    crates/tako/src/internal/server/reactor.rs:500-580     (DAG mode)
// It is NOT a port: the reference solves a MILP over (worker, class, variant) counts with HiGHS; this
// library runs a deterministic priority-ordered first-fit over the same aggregation, with the per-task
// work (histogram, stable ranking, emission) as streaming kernels over an SoA task table in HBM.
// See DESIGN.md for the data layout, the kernels and their rooflines.
//
// Device data (all SoA, indexed by dense task handle h):
//   key[h]   u32  bit31 READY | bit30 DONE (assigned by a tick) | bit29 VALID | level(15) | class(14)
//   prio[h]  u64  tako Priority (only read when the level table changes)
//   deps[h]  u32  unfinished dependencies (DAG mode), cons_off/cons: CSR of consumers
// One tick = 3 kernels on one stream:
//   count_k : per-chunk histogram of ready tasks by group g = level*Q + class   (HBM streaming, 4 B/task)
//   solve_k : CTA 0: sequential first-fit over non-empty groups, one thread per worker (worker state in
//             registers; amounts gcd-scaled to 32 bits when the tick allows it, else 64-bit);
//             CTAs 1..: exclusive scan of the per-chunk histograms over chunks (runs concurrently), then they
//             stand by: when CTA 0 finds the first saturated priority level it hands every worker to one warp
//             of those CTAs, which fills it independently (pack_body) and reports back;
//   emit_k  : stable rank of every ready task inside its group, rank -> (worker, variant) through
//             the solver's count segments, compact write of 8-byte assignments, READY -> DONE
// Sharded over several GPUs (one context per GPU, tasks block-sharded, workers replicated) a fourth kernel,
// xchg_k, stores the rank's count vector into every peer's exchange buffer over NVLink between count_k and
// solve_k; solve_k acquires the peers' flags and sums the vectors itself (no host collective).
// The algorithm has a sequential specification, tests/greedy_model.py, which the kernels equal bit for bit.
#include "../../include/hqsched.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <numeric>
#include <vector>

namespace {

typedef uint32_t u32;
typedef uint64_t u64;

constexpr u32 KEY_READY = 1u << 31;
constexpr u32 KEY_DONE = 1u << 30;
constexpr u32 KEY_VALID = 1u << 29;
constexpr u32 KEY_LEVEL_SHIFT = 14;
constexpr u32 KEY_LEVEL_MASK = 0x7FFFu;
constexpr u32 KEY_CLASS_MASK = 0x3FFFu;

constexpr u32 COUNT_THREADS = 1024;       // count_k: one uint4 (4 tasks) per thread and pass
constexpr u32 EMIT_SMEM_BUDGET = 96 * 1024;
constexpr u32 SEG_CAP = 1u << 20;         // (group, worker, variant) count segments per tick
constexpr u32 NEWPRIO_CAP = 4096;

__device__ __forceinline__ u32 key_level(u32 k) { return (k >> KEY_LEVEL_SHIFT) & KEY_LEVEL_MASK; }
__device__ __forceinline__ u32 key_class(u32 k) { return k & KEY_CLASS_MASK; }

// Per-group result of the solver, read by emit_k with one 16-byte load.
struct __align__(16) GroupOut {
    u32 k;        // tasks of this group assigned this tick (global count in sharded mode)
    u32 out_off;  // offset of the group's first assignment in the (local) output
    u32 seg_lo;   // first count segment
    u32 seg_n;    // number of count segments
};

struct TickHeaderOut {
    u32 n_assigned;  // local assignments
    u32 n_groups;
    u32 n_segments;
    u32 error;       // 1 = segment overflow, 2 = solver grid synchronisation timed out
    unsigned long long dbg[8];   // clock64 phase stamps of CTA 0 (debug)
};

// ------------------------------------------------------------------------------------------------
// level lookup: levels[] sorted by DESCENDING priority.  exact mode: index of the entry equal to p
// (or ~0u if absent); coarse mode: levels[i] is the lowest priority of bucket i, index of the first
// bucket whose bound <= p (clamped to the last bucket).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 find_level(const u64* __restrict__ levels, u32 n_levels, u64 p, bool coarse) {
    u32 lo = 0, hi = n_levels;  // first index with levels[i] <= p
    while (lo < hi) {
        u32 mid = (lo + hi) >> 1;
        if (__ldg(levels + mid) <= p) hi = mid; else lo = mid + 1;
    }
    if (coarse) return lo < n_levels ? lo : n_levels - 1;
    if (lo < n_levels && __ldg(levels + lo) == p) return lo;
    return ~0u;
}

// ready-set maintenance ---------------------------------------------------------------------------
__global__ void push_k(u32 n, const u32* __restrict__ task, const u32* __restrict__ cls,
                       const u64* __restrict__ prio_in, u32* __restrict__ key, u64* __restrict__ prio,
                       const u64* __restrict__ levels, u32 n_levels, int coarse, u32* __restrict__ newcnt,
                       u64* __restrict__ newprio) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    const u32 h = live ? task[i] : 0u;
    const u64 p = live ? prio_in[i] : 0ull;
    u32 lvl = (live && n_levels) ? find_level(levels, n_levels, p, coarse != 0) : ~0u;
    // unknown priority: report it to the host (one atomic per warp), key it provisionally to level 0;
    // relevel_k fixes every key once the host has merged the new priority
    const bool fresh = live && lvl == ~0u;
    const u32 fm = __ballot_sync(0xffffffffu, fresh);
    if (fm) {
        const u32 lane = threadIdx.x & 31;
        u32 slot0 = 0;
        if (lane == (u32)(__ffs(fm) - 1)) slot0 = atomicAdd(newcnt, (u32)__popc(fm));
        slot0 = __shfl_sync(0xffffffffu, slot0, __ffs(fm) - 1);
        if (fresh) {
            const u32 slot = slot0 + __popc(fm & ((1u << lane) - 1));
            if (slot < NEWPRIO_CAP) newprio[slot] = p;
            lvl = 0;
        }
    }
    if (!live) return;
    prio[h] = p;
    key[h] = KEY_READY | KEY_VALID | (lvl << KEY_LEVEL_SHIFT) | (cls[i] & KEY_CLASS_MASK);
}

__global__ void relevel_k(u32 n_handles, u32* __restrict__ key, const u64* __restrict__ prio,
                          const u64* __restrict__ levels, u32 n_levels, int coarse) {
    u32 h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n_handles) return;
    u32 k = key[h];
    if (!(k & KEY_VALID)) return;
    u32 lvl = find_level(levels, n_levels, prio[h], coarse != 0);
    if (lvl == ~0u) lvl = 0;
    key[h] = (k & ~(KEY_LEVEL_MASK << KEY_LEVEL_SHIFT)) | (lvl << KEY_LEVEL_SHIFT);
}

__global__ void remove_k(u32 n, const u32* __restrict__ task, u32* __restrict__ key, u32 n_handles) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 h = task[i];
    if (h < n_handles) key[h] &= ~KEY_READY;
}

__global__ void rearm_k(u32 n_handles, u32* __restrict__ key) {
    u32 h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n_handles) return;
    u32 k = key[h];
    if (k & KEY_DONE) key[h] = (k & ~KEY_DONE) | KEY_READY;
}

__global__ void dag_init_k(u32 n, const u32* __restrict__ cls, const u64* __restrict__ prio,
                           const u32* __restrict__ deps, u32* __restrict__ key,
                           const u64* __restrict__ levels, u32 n_levels, int coarse) {
    u32 h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n) return;
    u32 lvl = find_level(levels, n_levels, prio[h], coarse != 0);
    if (lvl == ~0u) lvl = 0;
    key[h] = KEY_VALID | (deps[h] == 0 ? KEY_READY : 0u) | (lvl << KEY_LEVEL_SHIFT) | (cls[h] & KEY_CLASS_MASK);
}

// task_finished (reactor.rs:545-571): one thread per (finished task, consumer) pair would need a
// segmented layout; out-degree is small (<= 8 in the benchmark DAG), so one thread per finished task.
__global__ void finished_k(u32 n, const u32* __restrict__ task, const u32* __restrict__ cons_off,
                           const u32* __restrict__ cons, u32* __restrict__ deps, u32* __restrict__ key,
                           u32* __restrict__ n_new) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    u32 made = 0;
    if (i < n) {
        u32 t = task[i];
        u32 lo = cons_off[t], hi = cons_off[t + 1];
        for (u32 e = lo; e < hi; ++e) {
            u32 c = cons[e];
            if (atomicSub(&deps[c], 1u) == 1u) {  // decrease_unfinished_deps() hit zero
                atomicOr(&key[c], KEY_READY);
                ++made;
            }
        }
    }
    made = __reduce_add_sync(0xffffffffu, made);
    if ((threadIdx.x & 31) == 0 && made) atomicAdd(n_new, made);
}

// ------------------------------------------------------------------------------------------------
// K1: count_k — histogram of ready tasks per group for one chunk of the task table.
// HBM traffic: 4 B read per table slot.  smem: G u32 counters.  One uint4 (4 tasks) per thread.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
count_k(const u32* __restrict__ key, u32 n_handles, u32 chunk, u32 Q, u32 G, u32* __restrict__ table,
        u32* __restrict__ total) {
    extern __shared__ u32 s_hist[];
    for (u32 g = threadIdx.x; g < G; g += blockDim.x) s_hist[g] = 0;
    __syncthreads();
    const u32 base = blockIdx.x * chunk;
    const u32 end = min(base + chunk, n_handles);
    // chunk and base are multiples of 256 => 16-byte aligned uint4 loads; a ragged tail is scalar.
    const u32 vec_end = base + ((end - base) & ~3u);
    for (u32 rowb = base; rowb < end; rowb += blockDim.x * 4) {
        const u32 i = rowb + threadIdx.x * 4;
        u32 k[4];
        if (i + 4 <= vec_end) {
            uint4 v = __ldg(reinterpret_cast<const uint4*>(key + i));
            k[0] = v.x; k[1] = v.y; k[2] = v.z; k[3] = v.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) k[j] = (i + j < end) ? __ldg(key + i + j) : 0u;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            // keys inside one warp are mostly distinct (levels x classes), so plain shared-memory atomics
            // beat warp aggregation (match.any costs one round per distinct key)
            if (k[j] & KEY_READY) atomicAdd(&s_hist[key_level(k[j]) * Q + key_class(k[j])], 1u);
        }
    }
    __syncthreads();
    u32* row = table + (size_t)blockIdx.x * G;
    for (u32 g = threadIdx.x; g < G; g += blockDim.x) {
        const u32 v = s_hist[g];
        row[g] = v;
        if (v) atomicAdd(&total[g], v);
    }
}

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

// ------------------------------------------------------------------------------------------------
// K3: emit_k — stable (handle-ordered) rank of every ready task inside its group, rank -> placement.
// Each warp owns a contiguous sub-chunk; per-warp group counters live in shared memory:
//   s_cnt[w][g]  first pass: tasks of group g in warp w's sub-chunk; then turned into the rank at which
//                warp w's first task of group g starts; second pass: running counter.
// The per-group solver output and (when they fit) the count segments are staged in shared memory.
// HBM traffic: 4 B read per table slot (second read hits L1/L2), 8 B written per assignment, 4 B key
// write-back per assignment.
// ------------------------------------------------------------------------------------------------
constexpr u32 EMIT_SEG_SMEM = 1024;
constexpr u32 EMIT_ROWS = 4;          // rows of 32 tasks per warp: chunk = warps * 128 task slots

// lanes of the warp holding the same group id, in constant time: one ballot per key bit (match.any
// iterates once per DISTINCT key, and a warp of 32 tasks holds ~30 distinct (level, class) keys)
__device__ __forceinline__ u32 same_key_lanes(u32 act, u32 g, u32 nbits) {
    u32 peers = act;
    for (u32 b = 0; b < nbits; ++b) {
        const u32 bit = (g >> b) & 1u;
        const u32 bal = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? bal : ~bal;
    }
    return peers;
}

__global__ void __launch_bounds__(1024)
emit_k(u32* __restrict__ key, u32 n_handles, u32 chunk, u32 Q, u32 G, u32 g_smem, u32 nbits, const u32* __restrict__ table,
       const u32* __restrict__ before, const GroupOut* __restrict__ gout, const u32* __restrict__ seg_cum,
       const u32* __restrict__ seg_wv, const TickHeaderOut* __restrict__ hdr, hqs_assignment* __restrict__ out,
       u32 out_cap) {
    extern __shared__ __align__(16) u32 s_emit[];
    const u32 nwarps = blockDim.x >> 5;
    u32* s_cnt = s_emit;                                              // [nwarps][G]
    GroupOut* s_go = reinterpret_cast<GroupOut*>(s_emit + nwarps * G);  // [G] when g_smem
    u32* s_segc = reinterpret_cast<u32*>(s_go + (g_smem ? G : 0));     // [EMIT_SEG_SMEM]
    u32* s_segw = s_segc + EMIT_SEG_SMEM;
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (u32 i = threadIdx.x; i < nwarps * G; i += blockDim.x) s_cnt[i] = 0;
    // A chunk holds an assigned task only if, for some group, fewer than k[g] tasks of the group precede
    // the chunk (the assigned ones are the first k[g] in handle order): in a drain tick only the first
    // chunks qualify, the rest leave after reading one table row.
    {
        const u32* row0 = table + (size_t)blockIdx.x * G;
        bool mine = false;
        for (u32 g = threadIdx.x; g < G; g += blockDim.x) {
            const u32 bef = before ? __ldg(before + g) : 0u;
            mine |= row0[g] + bef < gout[g].k;
        }
        if (!__syncthreads_or(mine)) return;
    }
    const u32 n_seg = hdr->n_segments;
    const bool seg_smem = n_seg <= EMIT_SEG_SMEM;
    if (g_smem)
        for (u32 g = threadIdx.x; g < G; g += blockDim.x) s_go[g] = gout[g];
    if (seg_smem)
        for (u32 i = threadIdx.x; i < n_seg; i += blockDim.x) { s_segc[i] = seg_cum[i]; s_segw[i] = seg_wv[i]; }
    __syncthreads();

    const u32 base = blockIdx.x * chunk;
    const u32 end = min(base + chunk, n_handles);
    // every warp owns EMIT_ROWS rows of 32 consecutive tasks; keys and peer masks stay in registers
    // between the counting pass and the emitting pass
    const u32 wbeg = base + warp * (32 * EMIT_ROWS);
    u32* mycnt = s_cnt + warp * G;
    u32 kk[EMIT_ROWS], gg[EMIT_ROWS], peers[EMIT_ROWS];
#pragma unroll
    for (int j = 0; j < (int)EMIT_ROWS; ++j) {
        const u32 i = wbeg + j * 32 + lane;
        kk[j] = i < end ? key[i] : 0u;
    }
#pragma unroll
    for (int j = 0; j < (int)EMIT_ROWS; ++j) {
        const bool ready = (kk[j] & KEY_READY) != 0;
        gg[j] = key_level(kk[j]) * Q + key_class(kk[j]);
        const u32 act = __ballot_sync(0xffffffffu, ready);
        peers[j] = same_key_lanes(act, gg[j], nbits);
        if (!ready) peers[j] = 0;
        // pass 1: per-warp counts (rows in order; the leader of each key adds its lanes)
        if (ready && (u32)(__ffs(peers[j]) - 1) == lane) mycnt[gg[j]] += __popc(peers[j]);
        __syncwarp();
    }
    __syncthreads();
    // turn counts into starting ranks: rank0(w, g) = table[b][g] + sum_{w' < w} cnt[w'][g]
    const u32* row = table + (size_t)blockIdx.x * G;
    for (u32 g = threadIdx.x; g < G; g += blockDim.x) {
        u32 run = row[g];
        for (u32 w2 = 0; w2 < nwarps; ++w2) {
            const u32 c = s_cnt[w2 * G + g];
            s_cnt[w2 * G + g] = run;
            run += c;
        }
    }
    __syncthreads();

    // pass 2: rank and emit
#pragma unroll
    for (int j = 0; j < (int)EMIT_ROWS; ++j) {
        const u32 i = wbeg + j * 32 + lane;
        const u32 k = kk[j], g = gg[j], pm = peers[j];
        if (pm) {
            const u32 leader = __ffs(pm) - 1;
            u32 r0 = 0;
            if (leader == lane) {
                r0 = mycnt[g];
                mycnt[g] = r0 + __popc(pm);
            }
            r0 = __shfl_sync(pm, r0, leader);
            const u32 r_loc = r0 + __popc(pm & ((1u << lane) - 1));      // rank among this rank's tasks
            const u32 bef = before ? __ldg(before + g) : 0u;
            const GroupOut go = g_smem ? s_go[g] : gout[g];
            if (r_loc + bef < go.k) {
                const u32 r = r_loc + bef;                                // global rank in the group
                // first segment whose inclusive end rank exceeds r
                u32 lo = go.seg_lo, hi = go.seg_lo + go.seg_n;
                u32 wv;
                if (seg_smem) {
                    while (lo < hi) {
                        const u32 mid = (lo + hi) >> 1;
                        if (s_segc[mid] > r) hi = mid; else lo = mid + 1;
                    }
                    wv = s_segw[lo];
                } else {
                    while (lo < hi) {
                        const u32 mid = (lo + hi) >> 1;
                        if (__ldg(seg_cum + mid) > r) hi = mid; else lo = mid + 1;
                    }
                    wv = __ldg(seg_wv + lo);
                }
                const u32 oi = go.out_off + r_loc;
                if (oi < out_cap) {
                    hqs_assignment asg;
                    asg.task = i;
                    asg.worker = (uint16_t)(wv & 0xFFFFu);
                    asg.variant = (uint8_t)(wv >> 16);
                    asg.kind = 0;
                    out[oi] = asg;
                }
                key[i] = (k & ~KEY_READY) | KEY_DONE;                     // Waiting -> Assigned
            }
        }
        __syncwarp();
    }
}

// per-worker totals of the count segments (what-if query)
__global__ void seg_worker_totals_k(const TickHeaderOut* __restrict__ hdr, const GroupOut* __restrict__ gout, u32 G,
                                    const u32* __restrict__ seg_cum, const u32* __restrict__ seg_wv, u32* __restrict__ per_worker) {
    // one thread per group: walks the group's segments (inclusive end ranks -> counts)
    const u32 g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    const GroupOut go = gout[g];
    if (go.k == 0) return;
    u32 prev = 0;
    for (u32 i = 0; i < go.seg_n; ++i) {
        const u32 end = seg_cum[go.seg_lo + i];
        atomicAdd(&per_worker[seg_wv[go.seg_lo + i] & 0xFFFFu], end - prev);
        prev = end;
    }
    (void)hdr;
}

// ================================================================================================
// host side
// ================================================================================================
thread_local std::string g_create_error;

}  // namespace

struct hqs_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    u32 R = 0;
    std::string err;
    // classes
    u32 Q = 0;
    std::vector<hqs_class> classes;
    unsigned char* d_classes = nullptr;   // ClassT<RT, u64>[Q]
    u32 d_classes_cap = 0;                // bytes
    unsigned char* d_classes32 = nullptr; // ClassT<RT, u32>[Q]: amounts / gscale[r] (valid when narrow_classes)
    u32 d_classes32_cap = 0;
    u32 class_bytes32 = 0;                // sizeof(ClassT<RT, u32>)
    u64 gscale[HQS_MAX_RESOURCES] = {};   // per-resource gcd of every requested amount (1 where nothing is requested)
    u64 narrow_limit[HQS_MAX_RESOURCES] = {};   // largest worker amount the narrow path can hold: gscale * (2^31 - 1)
    bool narrow_classes = false;          // every scaled class amount < 2^31
    // peer-to-peer sharded tick
    u32* d_xbuf = nullptr;                // my exchange buffer: [2][HQS_MAX_PEERS][HQS_MAX_GROUPS] counts + [2][HQS_MAX_PEERS] flags
    u32* x_peer[HQS_MAX_PEERS] = {};      // every rank's exchange buffer (own included), set by hqs_shard_attach
    std::vector<void*> x_opened;          // IPC mappings to close
    u32 x_world = 0, x_rank = 0, x_seq = 0;
    u32* d_xall = nullptr;                // [HQS_MAX_GROUPS] sum over ranks, written by the solver
    u32* d_xbefore = nullptr;             // [HQS_MAX_GROUPS] sum over lower ranks
    bool x_tick = false;                  // the tick being launched uses the exchange
    bool tick_narrow = false;             // this tick runs the narrow solver
    bool force_wide = false;              // hqs_create flag bit 1: always the 64-bit solver (tests)
    u32 RT = 4;                           // resource slots of the device class layout (4, 8 or 16)
    u32 class_bytes = 0;                  // sizeof(ClassT<RT>)
    // priority levels (descending)
    std::vector<u64> levels;      // exact distinct priorities seen, descending
    std::vector<u64> dev_levels;  // what the device uses (== levels, or bucket bounds when coarsened)
    bool coarse = false;
    u64* d_levels = nullptr;
    u32 d_levels_cap = 0;
    // task table
    u32 n_handles = 0, cap_handles = 0;
    u32* d_key = nullptr;
    u64* d_prio = nullptr;
    u32* d_deps = nullptr;
    u32* d_cons_off = nullptr;
    u32* d_cons = nullptr;
    bool dag = false;
    // push staging (device)
    u32* d_push_task = nullptr; u32* d_push_cls = nullptr; u64* d_push_prio = nullptr;
    u32 push_cap = 0;
    u32* d_newcnt = nullptr; u64* d_newprio = nullptr;
    // tick buffers
    u32 sm_count = 148;
    u32 G_cap = 0, P_cap = 0;
    u32* d_table = nullptr;
    u32* d_total = nullptr;
    GroupOut* d_gout = nullptr;
    uint2* d_glist = nullptr;
    SolveSync* d_sync = nullptr;
    u64* d_pk_fr = nullptr; u32* d_pk_quota = nullptr; u32* d_pk_taken = nullptr; u32* d_pk_cand = nullptr; u32* d_pk_meta = nullptr;
    u32* d_seg_cum = nullptr; u32* d_seg_wv = nullptr;
    hqs_assignment* d_out = nullptr; u32 out_cap_dev = 0;
    TickHeaderOut* d_hdr = nullptr;
    u64* d_free_after = nullptr;
    unsigned char* d_tickin = nullptr; size_t tickin_cap = 0;
    unsigned char* h_tickin = nullptr;  // pinned
    unsigned char* h_hdr = nullptr;     // pinned: TickHeaderOut + free_after
    size_t h_hdr_cap = 0;
    u32* h_small = nullptr;             // pinned scratch (counters)
    // last tick
    u32 last_W = 0, last_G = 0, last_L = 0;
    bool tick_pending = false;
    bool own_stream = true;
    bool profile = false;
    bool pack = true;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool ev_valid = false;
    hqs_stats stats{};
    unsigned long long dbg[8] = {0};
};

namespace {

int fail(hqs_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, HQS_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),   \
                        __FILE__, __LINE__);                                                       \
    } while (0)

template <typename T>
int dev_realloc(hqs_ctx* ctx, T** p, size_t old_n, size_t new_n, bool keep, bool zero_new) {
    T* q = nullptr;
    CU(cudaMalloc(&q, new_n * sizeof(T)));
    if (zero_new) CU(cudaMemsetAsync(q, 0, new_n * sizeof(T), ctx->stream));
    if (keep && *p && old_n) CU(cudaMemcpyAsync(q, *p, old_n * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
    if (*p) {
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaFree(*p));
    }
    *p = q;
    return HQS_OK;
}

int ensure_handles(hqs_ctx* ctx, u32 need) {
    if (need <= ctx->cap_handles) return HQS_OK;
    u32 cap = std::max<u32>(need, std::max<u32>(ctx->cap_handles * 2, 1u << 16));
    cap = (cap + 1023u) & ~1023u;
    int rc;
    if ((rc = dev_realloc(ctx, &ctx->d_key, ctx->cap_handles, cap, true, true))) return rc;
    if ((rc = dev_realloc(ctx, &ctx->d_prio, ctx->cap_handles, cap, true, true))) return rc;
    ctx->cap_handles = cap;
    return HQS_OK;
}

// (Re)builds the device level table from ctx->levels, coarsening when L * Q exceeds HQS_MAX_GROUPS.
int upload_levels(hqs_ctx* ctx) {
    const u32 q = std::max<u32>(ctx->Q, 1);
    const u32 max_levels = std::max<u32>(1, HQS_MAX_GROUPS / q);
    const u32 L = (u32)ctx->levels.size();
    ctx->dev_levels.clear();
    if (L <= max_levels) {
        ctx->dev_levels = ctx->levels;
        ctx->coarse = false;
    } else {
        // merge adjacent levels into max_levels buckets; entry i = lowest priority of bucket i
        ctx->coarse = true;
        for (u32 b = 0; b < max_levels; ++b) {
            const u64 last = ((u64)(b + 1) * L) / max_levels - 1;
            ctx->dev_levels.push_back(ctx->levels[last]);
        }
        ctx->dev_levels.back() = 0;  // the last bucket takes everything below
    }
    const u32 n = (u32)ctx->dev_levels.size();
    if (n > ctx->d_levels_cap) {
        if (ctx->d_levels) { CU(cudaStreamSynchronize(ctx->stream)); CU(cudaFree(ctx->d_levels)); ctx->d_levels = nullptr; }
        ctx->d_levels_cap = std::max<u32>(n * 2, 64);
        CU(cudaMalloc(&ctx->d_levels, ctx->d_levels_cap * sizeof(u64)));
    }
    if (n) {
        // pageable source: the copy is staged by the runtime before the call returns
        CU(cudaMemcpyAsync(ctx->d_levels, ctx->dev_levels.data(), n * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    ctx->stats.coarsened = ctx->coarse ? 1 : 0;
    return HQS_OK;
}

int relevel_all(hqs_ctx* ctx) {
    if (!ctx->n_handles || ctx->dev_levels.empty()) return HQS_OK;
    relevel_k<<<(ctx->n_handles + 255) / 256, 256, 0, ctx->stream>>>(
        ctx->n_handles, ctx->d_key, ctx->d_prio, ctx->d_levels, (u32)ctx->dev_levels.size(), ctx->coarse ? 1 : 0);
    ctx->stats.kernel_launches++;
    CU(cudaGetLastError());
    return HQS_OK;
}

// merges new distinct priorities into the level set; returns true if the set changed
bool merge_levels(hqs_ctx* ctx, std::vector<u64>& fresh) {
    std::sort(fresh.begin(), fresh.end(), std::greater<u64>());
    fresh.erase(std::unique(fresh.begin(), fresh.end()), fresh.end());
    std::vector<u64> merged;
    merged.reserve(ctx->levels.size() + fresh.size());
    std::merge(ctx->levels.begin(), ctx->levels.end(), fresh.begin(), fresh.end(), std::back_inserter(merged),
               std::greater<u64>());
    merged.erase(std::unique(merged.begin(), merged.end()), merged.end());
    const bool changed = merged.size() != ctx->levels.size();
    ctx->levels.swap(merged);
    return changed;
}

void distinct_priorities(const u64* p, u32 n, std::vector<u64>& out) {
    // small open-addressing set with a last-value fast path; distinct priorities are few
    std::vector<u64> slots(1024, 0);
    std::vector<unsigned char> used(1024, 0);
    size_t count = 0;
    u64 last = n ? ~p[0] : 0;
    for (u32 i = 0; i < n; ++i) {
        const u64 v = p[i];
        if (v == last) continue;
        last = v;
        if ((count + 1) * 2 > slots.size()) {
            std::vector<u64> ns(slots.size() * 4, 0);
            std::vector<unsigned char> nu(slots.size() * 4, 0);
            for (size_t s = 0; s < slots.size(); ++s)
                if (used[s]) {
                    size_t h = (slots[s] * 0x9E3779B97F4A7C15ull) >> 20 & (ns.size() - 1);
                    while (nu[h]) h = (h + 1) & (ns.size() - 1);
                    ns[h] = slots[s]; nu[h] = 1;
                }
            slots.swap(ns); used.swap(nu);
        }
        size_t h = (v * 0x9E3779B97F4A7C15ull) >> 20 & (slots.size() - 1);
        while (used[h] && slots[h] != v) h = (h + 1) & (slots.size() - 1);
        if (!used[h]) { used[h] = 1; slots[h] = v; ++count; }
    }
    for (size_t s = 0; s < slots.size(); ++s) if (used[s]) out.push_back(slots[s]);
}

int ensure_tick_buffers(hqs_ctx* ctx, u32 G, u32 P, u32 W, u32 out_cap) {
    if (G > ctx->G_cap || P > ctx->P_cap) {
        const u32 ng = std::max(G, ctx->G_cap), np = std::max(P, ctx->P_cap);
        CU(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_table) CU(cudaFree(ctx->d_table));
        CU(cudaMalloc(&ctx->d_table, (size_t)ng * np * sizeof(u32)));
        if (ng > ctx->G_cap) {
            if (ctx->d_total) CU(cudaFree(ctx->d_total));
            if (ctx->d_gout) CU(cudaFree(ctx->d_gout));
            if (ctx->d_glist) CU(cudaFree(ctx->d_glist));
            CU(cudaMalloc(&ctx->d_total, ng * sizeof(u32)));
            CU(cudaMemsetAsync(ctx->d_total, 0, ng * sizeof(u32), ctx->stream));
            CU(cudaMalloc(&ctx->d_gout, ng * sizeof(GroupOut)));
            CU(cudaMemsetAsync(ctx->d_gout, 0, ng * sizeof(GroupOut), ctx->stream));
            CU(cudaMalloc(&ctx->d_glist, ng * sizeof(uint2)));
        }
        ctx->G_cap = ng; ctx->P_cap = np;
    }
    if (!ctx->d_seg_cum) {
        CU(cudaMalloc(&ctx->d_seg_cum, SEG_CAP * sizeof(u32)));
        CU(cudaMalloc(&ctx->d_seg_wv, SEG_CAP * sizeof(u32)));
        CU(cudaMalloc(&ctx->d_hdr, sizeof(TickHeaderOut)));
        CU(cudaMalloc(&ctx->d_free_after, (size_t)HQS_MAX_WORKERS * HQS_MAX_RESOURCES * sizeof(u64)));
        CU(cudaMalloc(&ctx->d_sync, sizeof(SolveSync)));
        CU(cudaMalloc(&ctx->d_pk_fr, (size_t)HQS_MAX_WORKERS * HQS_MAX_RESOURCES * sizeof(u64)));
        CU(cudaMalloc(&ctx->d_pk_quota, (size_t)HQS_MAX_WORKERS * PACK_MAX_CAND * sizeof(u32)));
        CU(cudaMalloc(&ctx->d_pk_taken, (size_t)HQS_MAX_WORKERS * PACK_MAX_CAND * sizeof(u32)));
        CU(cudaMalloc(&ctx->d_pk_cand, PACK_MAX_CAND * sizeof(u32)));
        CU(cudaMalloc(&ctx->d_pk_meta, 2 * sizeof(u32)));
    }
    if (out_cap > ctx->out_cap_dev) {
        CU(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_out) CU(cudaFree(ctx->d_out));
        ctx->out_cap_dev = std::max<u32>(out_cap, 1024);
        CU(cudaMalloc(&ctx->d_out, (size_t)ctx->out_cap_dev * sizeof(hqs_assignment)));
    }
    (void)W;
    return HQS_OK;
}

struct TickLayout {
    size_t off_free, off_total, off_rem, off_order, off_vorder, off_blocked, bytes;
};

TickLayout tick_layout(u32 W, u32 R, u32 Q, bool blocked) {
    TickLayout l;
    size_t o = 0;
    l.off_free = o; o += (size_t)W * R * 8;
    l.off_total = o; o += (size_t)W * R * 8;
    l.off_rem = o; o += (size_t)W * 8;
    l.off_order = o; o += (size_t)Q * 4;
    l.off_vorder = o; o += (size_t)Q * HQS_MAX_VARIANTS;
    o = (o + 15) & ~size_t(15);
    l.off_blocked = o; if (blocked) o += (size_t)W * Q;
    l.bytes = (o + 15) & ~size_t(15);
    return l;
}

// Per-tick orders, both from S_r = sum over workers of free[r] (MAX counts as one unit):
//  order[]   classes inside one priority level by descending objective weight of one task, the greedy
//            analogue of the MILP coefficient of create_sn_var (solver.rs:520-549):
//            weight * sum_r amount_r / S_r
//  vorder[]  variants of a class by ascending dominant share max_r amount_r / S_r: the variant that costs
//            least of the scarcest thing it touches is tried first
void tick_orders(const hqs_ctx* ctx, u32 W, const u64* free_rw, const u64* total_rw, u32* order, uint8_t* vorder) {
    const u32 R = ctx->R, Q = ctx->Q;
    double S[HQS_MAX_RESOURCES], T[HQS_MAX_RESOURCES];
    for (u32 r = 0; r < R; ++r) { S[r] = 0; T[r] = 0; }
    for (u32 w = 0; w < W; ++w)
        for (u32 r = 0; r < R; ++r) {
            const u64 f = free_rw[(size_t)w * R + r];
            S[r] += f == HQS_AMOUNT_MAX ? 1.0 : (double)f / 10000.0;
            const u64 t = total_rw[(size_t)w * R + r];
            T[r] += t == HQS_AMOUNT_MAX ? 1.0 : (double)t / 10000.0;
        }
    std::vector<std::pair<double, u32>> sc(Q);
    for (u32 c = 0; c < Q; ++c) {
        double best = 0;
        const hqs_class& cl = ctx->classes[c];
        std::pair<double, u32> doms[HQS_MAX_VARIANTS];
        for (u32 v = 0; v < cl.n_variants; ++v) {
            double s = 0, dom = 0;
            for (u32 r = 0; r < R; ++r) {
                const bool all = (cl.variants[v].all_mask >> r) & 1;
                const u64 amt = all ? 0 : cl.variants[v].amount[r];
                if (amt) {
                    const double x = S[r] < 1e-6 ? INFINITY : ((double)amt / 10000.0) / S[r];
                    dom = x > dom ? x : dom;
                }
                if (S[r] < 1e-6) continue;
                if (all) s += (T[r] / std::max<u32>(W, 1)) / S[r];
                else s += ((double)amt / 10000.0) / S[r];
            }
            if (cl.variants[v].all_mask) dom = INFINITY;
            s *= (double)cl.variants[v].weight / 10000.0;
            best = std::max(best, s);
            doms[v] = {dom, v};
        }
        sc[c] = {best, c};
        std::stable_sort(doms, doms + cl.n_variants,
                         [](const std::pair<double, u32>& a, const std::pair<double, u32>& b) { return a.first < b.first; });
        for (u32 v = 0; v < HQS_MAX_VARIANTS; ++v) vorder[c * HQS_MAX_VARIANTS + v] = v < cl.n_variants ? (uint8_t)doms[v].second : 0;
    }
    std::stable_sort(sc.begin(), sc.end(), [](const std::pair<double, u32>& a, const std::pair<double, u32>& b) {
        return a.first > b.first;
    });
    for (u32 c = 0; c < Q; ++c) order[c] = sc[c].second;
}

struct TickGeom { u32 G, L, P, chunk, emit_warps, g_smem, nbits; size_t emit_smem; };

TickGeom tick_geom(const hqs_ctx* ctx) {
    TickGeom t;
    t.L = std::max<u32>((u32)ctx->dev_levels.size(), 1);
    t.G = t.L * std::max<u32>(ctx->Q, 1);
    t.nbits = 1;
    while ((1u << t.nbits) < t.G) t.nbits++;
    // emit_k shared memory: warps * G counters (+ G solver records + the segment cache); as many warps per
    // CTA as the budget allows, two CTAs per SM
    const size_t seg_cache = 2 * EMIT_SEG_SMEM * sizeof(u32);
    t.emit_warps = 8;
    for (u32 ew : {32u, 16u}) {
        if ((size_t)ew * t.G * 4 + (size_t)t.G * sizeof(GroupOut) + seg_cache <= 64 * 1024) { t.emit_warps = ew; break; }
    }
    t.g_smem = ((size_t)t.emit_warps * t.G * 4 + (size_t)t.G * sizeof(GroupOut) + seg_cache <= EMIT_SMEM_BUDGET) ? 1 : 0;
    t.emit_smem = (size_t)t.emit_warps * t.G * 4 + (t.g_smem ? (size_t)t.G * sizeof(GroupOut) : 0) + seg_cache;
    const u32 n = std::max<u32>(ctx->n_handles, 1);
    const u32 chunk = t.emit_warps * 32 * EMIT_ROWS;     // every emit warp owns EMIT_ROWS rows
    t.chunk = chunk;
    t.P = (n + chunk - 1) / chunk;
    return t;
}

int upload_tick_input(hqs_ctx* ctx, u32 W, const hqs_worker* workers, const u64* free_rw, const u64* total_rw,
                      const uint8_t* blocked, TickLayout* lay_out) {
    const u32 R = ctx->R, Q = ctx->Q;
    const TickLayout lay = tick_layout(W, R, Q, blocked != nullptr);
    if (lay.bytes > ctx->tickin_cap) {
        CU(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_tickin) CU(cudaFree(ctx->d_tickin));
        if (ctx->h_tickin) CU(cudaFreeHost(ctx->h_tickin));
        ctx->tickin_cap = lay.bytes * 2;
        CU(cudaMalloc(&ctx->d_tickin, ctx->tickin_cap));
        CU(cudaMallocHost(&ctx->h_tickin, ctx->tickin_cap));
    }
    unsigned char* h = ctx->h_tickin;
    memcpy(h + lay.off_free, free_rw, (size_t)W * R * 8);
    memcpy(h + lay.off_total, total_rw, (size_t)W * R * 8);
    // narrow solver: every worker amount of this tick must fit 31 bits after the per-resource scaling
    static const bool force_wide = getenv("HQS_DEBUG_WIDE") != nullptr;
    bool narrow = ctx->narrow_classes && !force_wide && !ctx->force_wide;
    if (narrow) {
        u64 over = 0;
        for (u32 w = 0; w < W; ++w)
            for (u32 r = 0; r < R; ++r) {
                const u64 f = free_rw[(size_t)w * R + r], t = total_rw[(size_t)w * R + r], lim = ctx->narrow_limit[r];
                over |= (u64)(f != HQS_AMOUNT_MAX && f > lim) | (u64)(t != HQS_AMOUNT_MAX && t > lim);
            }
        narrow = over == 0;
    }
    ctx->tick_narrow = narrow;
    ctx->stats.narrow_amounts = narrow ? 1 : 0;
    u64* rem = reinterpret_cast<u64*>(h + lay.off_rem);
    for (u32 w = 0; w < W; ++w) rem[w] = workers[w].remaining_time_ms;
    tick_orders(ctx, W, free_rw, total_rw, reinterpret_cast<u32*>(h + lay.off_order), h + lay.off_vorder);
    if (blocked) {
        // ABI bit index ((w*Q + c) * HQS_MAX_VARIANTS + v) with HQS_MAX_VARIANTS == 8: one byte per (w, c)
        memcpy(h + lay.off_blocked, blocked, (size_t)W * Q);
    }
    CU(cudaMemcpyAsync(ctx->d_tickin, h, lay.bytes, cudaMemcpyHostToDevice, ctx->stream));
    *lay_out = lay;
    return HQS_OK;
}

// maxima of two u32 arrays, eight independent accumulators each (the compiler turns them into vector max)
void max_of_u32_pair(const u32* a, const u32* b, u32 n, u32* max_a, u32* max_b) {
    u32 ma[8] = {0, 0, 0, 0, 0, 0, 0, 0}, mb[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    u32 i = 0;
    for (; i + 8 <= n; i += 8)
        for (u32 j = 0; j < 8; ++j) {
            ma[j] = a[i + j] > ma[j] ? a[i + j] : ma[j];
            mb[j] = b[i + j] > mb[j] ? b[i + j] : mb[j];
        }
    for (; i < n; ++i) {
        ma[0] = a[i] > ma[0] ? a[i] : ma[0];
        mb[0] = b[i] > mb[0] ? b[i] : mb[0];
    }
    for (u32 j = 1; j < 8; ++j) { ma[0] = ma[j] > ma[0] ? ma[j] : ma[0]; mb[0] = mb[j] > mb[0] ? mb[j] : mb[0]; }
    *max_a = ma[0];
    *max_b = mb[0];
}

int validate_workers(hqs_ctx* ctx, u32 W, const hqs_worker* workers, const u64* free_rw, const u64* total_rw) {
    if (!workers || !free_rw || !total_rw) return fail(ctx, HQS_E_INVALID, "null worker arrays");
    if (W == 0 || W > HQS_MAX_WORKERS) return fail(ctx, HQS_E_LIMIT, "n_workers=%u outside 1..%u", W, HQS_MAX_WORKERS);
    for (u32 w = 1; w < W; ++w)
        if (workers[w].worker_id <= workers[w - 1].worker_id)
            return fail(ctx, HQS_E_INVALID, "workers must be sorted by ascending unique worker_id");
    if (ctx->Q == 0) return fail(ctx, HQS_E_STATE, "hqs_classes_set has not been called");
    return HQS_OK;
}

int launch_count(hqs_ctx* ctx, const TickGeom& t) {
    ctx->ev_valid = false;
    if (ctx->profile) CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    count_k<<<t.P, COUNT_THREADS, t.G * sizeof(u32), ctx->stream>>>(ctx->d_key, ctx->n_handles, t.chunk, ctx->Q, t.G,
                                                                   ctx->d_table, ctx->d_total);
    ctx->stats.kernel_launches++;
    CU(cudaGetLastError());
    if (ctx->profile) CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    return HQS_OK;
}

int launch_solve_emit(hqs_ctx* ctx, const TickGeom& t, u32 W, const TickLayout& lay, bool blocked,
                      const u32* d_counts_all, const u32* d_before, u32 out_cap, bool no_emit = false) {
    SolveArgs a;
    a.free_rw = reinterpret_cast<const u64*>(ctx->d_tickin + lay.off_free);
    a.total_rw = reinterpret_cast<const u64*>(ctx->d_tickin + lay.off_total);
    a.rem_time = reinterpret_cast<const u64*>(ctx->d_tickin + lay.off_rem);
    a.order = reinterpret_cast<const u32*>(ctx->d_tickin + lay.off_order);
    a.vorder = ctx->d_tickin + lay.off_vorder;
    a.blocked = blocked ? ctx->d_tickin + lay.off_blocked : nullptr;
    const bool narrow = ctx->tick_narrow;
    a.classes = narrow ? ctx->d_classes32 : ctx->d_classes;
    a.classes64 = ctx->d_classes;
    for (u32 r = 0; r < HQS_MAX_RESOURCES; ++r) a.gscale[r] = ctx->gscale[r] ? ctx->gscale[r] : 1;
    a.W = W; a.Q = ctx->Q; a.L = t.L; a.R = ctx->R; a.G = t.G;
    a.classes_bytes = ctx->Q * (narrow ? ctx->class_bytes32 : ctx->class_bytes);
    a.total_local = ctx->d_total;
    a.total_all = d_counts_all ? d_counts_all : ctx->d_total;
    a.before = d_before;
    a.gout = ctx->d_gout;
    a.seg_cum = ctx->d_seg_cum; a.seg_wv = ctx->d_seg_wv;
    a.free_after = ctx->d_free_after;
    a.hdr = ctx->d_hdr;
    a.glist = ctx->d_glist;
    a.table = ctx->d_table; a.P = t.P;
    a.x_world = 0; a.x_rank = 0; a.x_seq = 0; a.x_counts = nullptr; a.x_flags = nullptr; a.x_all = nullptr; a.x_before = nullptr;
    if (ctx->x_tick) {
        const u32 parity = ctx->x_seq & 1u;
        a.x_world = ctx->x_world; a.x_rank = ctx->x_rank; a.x_seq = ctx->x_seq;
        a.x_counts = ctx->d_xbuf + (size_t)parity * HQS_MAX_PEERS * HQS_MAX_GROUPS;
        a.x_flags = ctx->d_xbuf + (size_t)2 * HQS_MAX_PEERS * HQS_MAX_GROUPS + (size_t)parity * HQS_MAX_PEERS;
        a.x_all = ctx->d_xall; a.x_before = ctx->d_xbefore;
    }
    a.sync = ctx->d_sync;
    a.pk.fr = ctx->d_pk_fr; a.pk.quota = ctx->d_pk_quota; a.pk.taken = ctx->d_pk_taken;
    a.pk.cand = ctx->d_pk_cand; a.pk.meta = ctx->d_pk_meta;
    const u32 threads = std::max<u32>(128, (W + 31) / 32 * 32);
    const u32 nw = threads / 32;
    a.scan_ctas = std::min<u32>((t.G + nw - 1) / nw, ctx->sm_count - 1);
    // grid: CTA 0 solves; the others scan the chunk table and then stand by to fill workers (one warp per
    // worker, spread over the SMs).  All CTAs must be co-resident (cooperative launch): <= one per SM.
    u32 grid = std::max<u32>(1 + a.scan_ctas, std::min<u32>(ctx->sm_count, 1 + (W + 1) / 2));
    grid = std::min<u32>(grid, ctx->sm_count);
    static const bool dbg_small_grid = getenv("HQS_DEBUG_SMALL_GRID") != nullptr;   // profiling aid: solver CTA + scan CTAs only
    if (dbg_small_grid) grid = 1 + a.scan_ctas;
    a.pack_enabled = (ctx->pack && grid >= 2 && !dbg_small_grid) ? 1 : 0;
    // SMALL variant: class table + variant order staged in shared memory
    a.smem_classes = ((size_t)a.classes_bytes + (size_t)ctx->Q * HQS_MAX_VARIANTS <= 48 * 1024) ? 1 : 0;
    a.smem_vorder = a.smem_classes;
    size_t solve_smem = 0;
    if (a.smem_classes) solve_smem += ((a.classes_bytes + 15u) & ~15u) + ((ctx->Q * HQS_MAX_VARIANTS + 15u) & ~15u);
    a.smem_glist_cap = std::min<u32>(t.G, 2048);
    solve_smem += (size_t)a.smem_glist_cap * (sizeof(uint2) + sizeof(u32) + 1) + (size_t)((a.smem_glist_cap + 31) / 32) * 4 + 32;
    solve_smem += (size_t)a.smem_glist_cap * sizeof(GroupOut) + 2 * SEG_SMEM * sizeof(u32);
    solve_smem = std::max(solve_smem, (size_t)nw * PACK_MAX_CAND * sizeof(double));     // pack warps' scratch
    CU(cudaMemsetAsync(ctx->d_sync, 0, sizeof(SolveSync), ctx->stream));
    void* kargs[] = {&a};
    const bool small = a.smem_classes != 0;
    const void* fn;
#define HQS_PICK(MT, SM, AT) (ctx->RT == 4 ? (const void*)solve_k<4, MT, SM, AT> : ctx->RT == 8 ? (const void*)solve_k<8, MT, SM, AT> : (const void*)solve_k<16, MT, SM, AT>)
#define HQS_PICK2(MT, SM) (narrow ? HQS_PICK(MT, SM, u32) : HQS_PICK(MT, SM, u64))
    if (threads <= 256) fn = small ? HQS_PICK2(256, true) : HQS_PICK2(256, false);
    else fn = small ? HQS_PICK2(1024, true) : HQS_PICK2(1024, false);
#undef HQS_PICK2
#undef HQS_PICK
    static const bool no_coop = getenv("HQS_DEBUG_NO_COOP") != nullptr;   // profiling aid: ncu skips cooperative launches
    if (no_coop) CU(cudaLaunchKernel(fn, dim3(grid), dim3(threads), kargs, solve_smem, ctx->stream));
    else CU(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(threads), kargs, solve_smem, ctx->stream));
    ctx->stats.kernel_launches++;
    if (ctx->profile) CU(cudaEventRecord(ctx->ev[2], ctx->stream));
    if (!no_emit)
    emit_k<<<t.P, 32 * t.emit_warps, t.emit_smem, ctx->stream>>>(
        ctx->d_key, ctx->n_handles, t.chunk, ctx->Q, t.G, t.g_smem, t.nbits, ctx->d_table, d_before, ctx->d_gout, ctx->d_seg_cum,
        ctx->d_seg_wv, ctx->d_hdr, ctx->d_out, out_cap);
    ctx->stats.kernel_launches++;
    CU(cudaGetLastError());
    if (ctx->profile) { CU(cudaEventRecord(ctx->ev[3], ctx->stream)); ctx->ev_valid = true; }
    ctx->last_W = W; ctx->last_G = t.G; ctx->last_L = t.L;
    ctx->tick_pending = true;
    ctx->stats.ticks++;
    return HQS_OK;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int hqs_abi_version(void) { return HQS_ABI_VERSION; }

const char* hqs_last_error(const hqs_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int hqs_create(hqs_ctx** out, int device, uint32_t n_resources, uint32_t flags) {
    hqs_ctx* ctx = nullptr;
    if (!out) return fail(nullptr, HQS_E_INVALID, "out is null");
    *out = nullptr;
    if (n_resources == 0 || n_resources > HQS_MAX_RESOURCES)
        return fail(nullptr, HQS_E_LIMIT, "n_resources=%u outside 1..%u", n_resources, HQS_MAX_RESOURCES);
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(nullptr, HQS_E_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                    cudaGetErrorString(e));
    if (device < 0 || device >= n_dev) return fail(nullptr, HQS_E_INVALID, "device %d out of range", device);
    ctx = new (std::nothrow) hqs_ctx();
    if (!ctx) return fail(nullptr, HQS_E_NOMEM, "out of memory");
    ctx->device = device;
    ctx->R = n_resources;
    ctx->RT = n_resources <= 4 ? 4 : n_resources <= 8 ? 8 : 16;
    ctx->class_bytes = ctx->RT == 4 ? sizeof(ClassT<4>) : ctx->RT == 8 ? sizeof(ClassT<8>) : sizeof(ClassT<16>);
    ctx->class_bytes32 = ctx->RT == 4 ? sizeof(ClassT<4, u32>) : ctx->RT == 8 ? sizeof(ClassT<8, u32>) : sizeof(ClassT<16, u32>);
    ctx->pack = !(flags & 1u);
    ctx->force_wide = (flags & 2u) != 0;
    e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    int sms = 0;
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_newcnt, sizeof(u32));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_newprio, NEWPRIO_CAP * sizeof(u64));
    if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_small, 64 * sizeof(u32));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(emit_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    {
        const void* fns[] = {
#define HQS_ALL(RTV, AT) (const void*)solve_k<RTV, 256, true, AT>, (const void*)solve_k<RTV, 256, false, AT>, \
                         (const void*)solve_k<RTV, 1024, true, AT>, (const void*)solve_k<RTV, 1024, false, AT>
            HQS_ALL(4, u64), HQS_ALL(8, u64), HQS_ALL(16, u64), HQS_ALL(4, u32), HQS_ALL(8, u32), HQS_ALL(16, u32)
#undef HQS_ALL
        };
        // dynamic + static shared memory of a CTA may not exceed 227 KB: allow each instance what its statics leave
        for (const void* f : fns) {
            cudaFuncAttributes fa;
            if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, f);
            if (e == cudaSuccess) {
                const size_t room = 227 * 1024 - fa.sharedSizeBytes;
                e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::min<size_t>(room, 200 * 1024));
            }
        }
    }
    if (e != cudaSuccess) {
        fail(nullptr, HQS_E_CUDA, "context setup failed: %s", cudaGetErrorString(e));
        delete ctx;
        return HQS_E_CUDA;
    }
    ctx->sm_count = sms > 0 ? (u32)sms : 148;
    *out = ctx;
    return HQS_OK;
}

void hqs_destroy(hqs_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    void* dev_ptrs[] = {ctx->d_classes, ctx->d_classes32, ctx->d_levels, ctx->d_key, ctx->d_prio, ctx->d_deps, ctx->d_cons_off,
                        ctx->d_cons, ctx->d_push_task, ctx->d_push_cls, ctx->d_push_prio, ctx->d_newcnt,
                        ctx->d_newprio, ctx->d_table, ctx->d_total, ctx->d_gout, ctx->d_glist, ctx->d_seg_cum,
                        ctx->d_seg_wv, ctx->d_out, ctx->d_hdr, ctx->d_free_after, ctx->d_tickin, ctx->d_sync, ctx->d_pk_fr,
                        ctx->d_pk_quota, ctx->d_pk_taken, ctx->d_pk_cand, ctx->d_pk_meta};
    for (void* p : dev_ptrs) if (p) cudaFree(p);
    for (void* p : ctx->x_opened) cudaIpcCloseMemHandle(p);
    if (ctx->d_xbuf) cudaFree(ctx->d_xbuf);
    if (ctx->d_xall) cudaFree(ctx->d_xall);
    if (ctx->d_xbefore) cudaFree(ctx->d_xbefore);
    if (ctx->h_tickin) cudaFreeHost(ctx->h_tickin);
    if (ctx->h_hdr) cudaFreeHost(ctx->h_hdr);
    if (ctx->h_small) cudaFreeHost(ctx->h_small);
    for (cudaEvent_t e : ctx->ev) if (e) cudaEventDestroy(e);
    if (ctx->stream && ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int hqs_classes_set(hqs_ctx* ctx, uint32_t n_classes, const hqs_class* classes) {
    if (!ctx) return HQS_E_INVALID;
    if (!classes || n_classes == 0) return fail(ctx, HQS_E_INVALID, "empty class table");
    if (n_classes > HQS_MAX_CLASSES) return fail(ctx, HQS_E_LIMIT, "n_classes=%u > %u", n_classes, HQS_MAX_CLASSES);
    // device layout: ClassT<RT>[Q] with RT = 4 / 8 / 16 resource slots, built as raw bytes
    const u32 RT = ctx->RT;
    const size_t var_bytes = (size_t)RT * 16 + 16, cls_bytes = ctx->class_bytes;
    std::vector<unsigned char> blob((size_t)n_classes * cls_bytes, 0);
    for (u32 c = 0; c < n_classes; ++c) {
        const hqs_class& sc = classes[c];
        if (sc.n_nodes != 0) return fail(ctx, HQS_E_INVALID, "class %u: multi-node requests are outside this path", c);
        if (sc.n_variants == 0 || sc.n_variants > HQS_MAX_VARIANTS)
            return fail(ctx, HQS_E_LIMIT, "class %u: n_variants=%u outside 1..%u", c, sc.n_variants, HQS_MAX_VARIANTS);
        unsigned char* cb = blob.data() + (size_t)c * cls_bytes;
        memcpy(cb, &sc.n_variants, 4);
        for (u32 v = 0; v < sc.n_variants; ++v) {
            unsigned char* vb = cb + 8 + (size_t)v * var_bytes;
            u64* amount = reinterpret_cast<u64*>(vb);
            float* rcp = reinterpret_cast<float*>(vb + (size_t)RT * 8);
            u64* min_time = reinterpret_cast<u64*>(vb + (size_t)RT * 16);
            u32* masks = reinterpret_cast<u32*>(vb + (size_t)RT * 16 + 8);
            u32 used = 0;
            for (u32 r = 0; r < HQS_MAX_RESOURCES; ++r) {
                const bool all = (sc.variants[v].all_mask >> r) & 1;
                const u64 amt = sc.variants[v].amount[r];
                if ((all || amt) && r >= ctx->R)
                    return fail(ctx, HQS_E_INVALID, "class %u variant %u uses resource %u >= n_resources", c, v, r);
                if (r < RT) {
                    amount[r] = all ? 0 : amt;
                    rcp[r] = (!all && amt) ? 1.0f / (float)amt : 0.0f;
                    rcp[RT + r] = all ? 0.0f : (float)(double)amt;          // u64 -> double -> float, both RN
                }
                if (all || amt) used |= 1u << r;
            }
            if (!used) return fail(ctx, HQS_E_INVALID, "class %u variant %u: empty request (request.rs:191-194)", c, v);
            if (sc.variants[v].weight == 0) return fail(ctx, HQS_E_INVALID, "class %u variant %u: zero weight", c, v);
            *min_time = sc.variants[v].min_time_ms;
            masks[0] = sc.variants[v].all_mask & ((1u << ctx->R) - 1);
            masks[1] = used;
        }
    }
    CU(cudaSetDevice(ctx->device));
    if (blob.size() > ctx->d_classes_cap) {
        CU(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_classes) CU(cudaFree(ctx->d_classes));
        ctx->d_classes_cap = (u32)std::max<size_t>(blob.size() * 2, 4096);
        CU(cudaMalloc(&ctx->d_classes, ctx->d_classes_cap));
    }
    CU(cudaMemcpyAsync(ctx->d_classes, blob.data(), blob.size(), cudaMemcpyHostToDevice, ctx->stream));
    // narrow copy: amounts divided by the per-resource gcd of everything requested
    u64 gs[HQS_MAX_RESOURCES];
    for (u32 r = 0; r < HQS_MAX_RESOURCES; ++r) gs[r] = 0;
    for (u32 c = 0; c < n_classes; ++c)
        for (u32 v = 0; v < classes[c].n_variants; ++v)
            for (u32 r = 0; r < ctx->R; ++r)
                if (!((classes[c].variants[v].all_mask >> r) & 1) && classes[c].variants[v].amount[r])
                    gs[r] = std::gcd(gs[r], (u64)classes[c].variants[v].amount[r]);
    bool narrow_ok = true;
    const size_t var_bytes32 = (size_t)RT * 12 + 16, cls_bytes32 = ctx->class_bytes32;
    std::vector<unsigned char> blob32((size_t)n_classes * cls_bytes32, 0);
    for (u32 r = 0; r < HQS_MAX_RESOURCES; ++r) {
        if (gs[r] == 0) gs[r] = 1;
        ctx->gscale[r] = gs[r];
        ctx->narrow_limit[r] = gs[r] > HQS_AMOUNT_MAX / NARROW_LIMIT ? HQS_AMOUNT_MAX - 1 : gs[r] * NARROW_LIMIT;
    }
    for (u32 c = 0; c < n_classes && narrow_ok; ++c) {
        const hqs_class& sc = classes[c];
        unsigned char* cb = blob32.data() + (size_t)c * cls_bytes32;
        memcpy(cb, &sc.n_variants, 4);
        for (u32 v = 0; v < sc.n_variants; ++v) {
            unsigned char* vb = cb + 8 + (size_t)v * var_bytes32;
            u32* amount = reinterpret_cast<u32*>(vb);
            float* rcp = reinterpret_cast<float*>(vb + (size_t)RT * 4);
            u64* min_time = reinterpret_cast<u64*>(vb + (size_t)RT * 12);
            u32* masks = reinterpret_cast<u32*>(vb + (size_t)RT * 12 + 8);
            u32 used = 0;
            for (u32 r = 0; r < ctx->R; ++r) {
                const bool all = (sc.variants[v].all_mask >> r) & 1;
                const u64 amt = all ? 0 : sc.variants[v].amount[r] / gs[r];
                if (amt > NARROW_LIMIT) narrow_ok = false;
                amount[r] = (u32)amt;
                rcp[r] = amt ? (1.0f / (float)(u32)amt) * (1.0f - 4.76837158203125e-7f) : 0.0f;   // biased low by 2^-21, see fit_count
                rcp[RT + r] = all ? 0.0f : (float)(double)sc.variants[v].amount[r];
                if (all || sc.variants[v].amount[r]) used |= 1u << r;
            }
            *min_time = sc.variants[v].min_time_ms;
            masks[0] = sc.variants[v].all_mask & ((1u << ctx->R) - 1);
            masks[1] = used;
        }
    }
    ctx->narrow_classes = narrow_ok;
    if (blob32.size() > ctx->d_classes32_cap) {
        CU(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_classes32) CU(cudaFree(ctx->d_classes32));
        ctx->d_classes32_cap = (u32)std::max<size_t>(blob32.size() * 2, 4096);
        CU(cudaMalloc(&ctx->d_classes32, ctx->d_classes32_cap));
    }
    CU(cudaMemcpyAsync(ctx->d_classes32, blob32.data(), blob32.size(), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    const bool q_changed = ctx->Q != n_classes;
    ctx->classes.assign(classes, classes + n_classes);
    ctx->Q = n_classes;
    if (q_changed && !ctx->levels.empty()) {
        // the level budget depends on Q: re-derive (possibly coarsened) levels and re-key the table
        const bool was_coarse = ctx->coarse;
        const size_t old_n = ctx->dev_levels.size();
        int rc = upload_levels(ctx);
        if (rc) return rc;
        if (was_coarse || ctx->coarse || old_n != ctx->dev_levels.size())
            if ((rc = relevel_all(ctx))) return rc;
    }
    return HQS_OK;
}

int hqs_ready_push(hqs_ctx* ctx, uint32_t n, const uint32_t* task, const uint32_t* class_id,
                   const uint64_t* priority) {
    if (!ctx) return HQS_E_INVALID;
    if (n == 0) return HQS_OK;
    if (!task || !class_id || !priority) return fail(ctx, HQS_E_INVALID, "null task arrays");
    if (ctx->Q == 0) return fail(ctx, HQS_E_STATE, "hqs_classes_set has not been called");
    if (ctx->dag) return fail(ctx, HQS_E_STATE, "hqs_ready_push is not available after hqs_dag_load");
    CU(cudaSetDevice(ctx->device));
    if (n > ctx->push_cap) {
        CU(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_push_task) { CU(cudaFree(ctx->d_push_task)); CU(cudaFree(ctx->d_push_cls)); CU(cudaFree(ctx->d_push_prio)); }
        ctx->push_cap = std::max<u32>(n, 1u << 16);
        CU(cudaMalloc(&ctx->d_push_task, (size_t)ctx->push_cap * 4));
        CU(cudaMalloc(&ctx->d_push_cls, (size_t)ctx->push_cap * 4));
        CU(cudaMalloc(&ctx->d_push_prio, (size_t)ctx->push_cap * 8));
    }
    CU(cudaMemcpyAsync(ctx->d_push_task, task, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_push_cls, class_id, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_push_prio, priority, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    // validate on the host while the staging copies are in flight (they touch no scheduler state)
    u32 max_h = 0, max_c = 0;
    max_of_u32_pair(task, class_id, n, &max_h, &max_c);
    if (max_c >= ctx->Q || max_h == ~0u) {
        cudaStreamSynchronize(ctx->stream);      // the caller may free its arrays as soon as we return
        if (max_c >= ctx->Q) return fail(ctx, HQS_E_INVALID, "class id %u >= n_classes %u", max_c, ctx->Q);
        return fail(ctx, HQS_E_INVALID, "task handle 0xFFFFFFFF is reserved");
    }
    int rc = ensure_handles(ctx, max_h + 1);
    if (rc) { cudaStreamSynchronize(ctx->stream); return rc; }
    ctx->n_handles = std::max(ctx->n_handles, max_h + 1);
    ctx->stats.n_handles = ctx->n_handles;
    CU(cudaMemsetAsync(ctx->d_newcnt, 0, sizeof(u32), ctx->stream));
    push_k<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, ctx->d_push_task, ctx->d_push_cls, ctx->d_push_prio, ctx->d_key,
                                                     ctx->d_prio, ctx->d_levels, (u32)ctx->dev_levels.size(),
                                                     ctx->coarse ? 1 : 0, ctx->d_newcnt, ctx->d_newprio);
    ctx->stats.kernel_launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(ctx->h_small, ctx->d_newcnt, sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    const u32 newcnt = ctx->h_small[0];
    if (newcnt) {
        std::vector<u64> fresh;
        if (newcnt <= NEWPRIO_CAP) {
            fresh.resize(newcnt);
            CU(cudaMemcpy(fresh.data(), ctx->d_newprio, newcnt * sizeof(u64), cudaMemcpyDeviceToHost));
        } else {
            distinct_priorities(priority, n, fresh);
        }
        merge_levels(ctx, fresh);
        if ((rc = upload_levels(ctx))) return rc;
        if ((rc = relevel_all(ctx))) return rc;
    }
    return HQS_OK;
}

int hqs_levels_add(hqs_ctx* ctx, uint32_t n, const uint64_t* priority) {
    if (!ctx) return HQS_E_INVALID;
    if (n == 0) return HQS_OK;
    if (!priority) return fail(ctx, HQS_E_INVALID, "null priority array");
    CU(cudaSetDevice(ctx->device));
    std::vector<u64> fresh;
    distinct_priorities(priority, n, fresh);
    if (!merge_levels(ctx, fresh)) return HQS_OK;
    int rc = upload_levels(ctx);
    if (rc) return rc;
    return relevel_all(ctx);
}

int hqs_ready_remove(hqs_ctx* ctx, uint32_t n, const uint32_t* task) {
    if (!ctx) return HQS_E_INVALID;
    if (n == 0) return HQS_OK;
    if (!task) return fail(ctx, HQS_E_INVALID, "null task array");
    CU(cudaSetDevice(ctx->device));
    if (n > ctx->push_cap) {
        CU(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_push_task) { CU(cudaFree(ctx->d_push_task)); CU(cudaFree(ctx->d_push_cls)); CU(cudaFree(ctx->d_push_prio)); }
        ctx->push_cap = std::max<u32>(n, 1u << 16);
        CU(cudaMalloc(&ctx->d_push_task, (size_t)ctx->push_cap * 4));
        CU(cudaMalloc(&ctx->d_push_cls, (size_t)ctx->push_cap * 4));
        CU(cudaMalloc(&ctx->d_push_prio, (size_t)ctx->push_cap * 8));
    }
    CU(cudaMemcpyAsync(ctx->d_push_task, task, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    remove_k<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, ctx->d_push_task, ctx->d_key, ctx->n_handles);
    ctx->stats.kernel_launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->stream));
    return HQS_OK;
}

int hqs_ready_rearm(hqs_ctx* ctx) {
    if (!ctx) return HQS_E_INVALID;
    if (!ctx->n_handles) return HQS_OK;
    CU(cudaSetDevice(ctx->device));
    rearm_k<<<(ctx->n_handles + 255) / 256, 256, 0, ctx->stream>>>(ctx->n_handles, ctx->d_key);
    ctx->stats.kernel_launches++;
    CU(cudaGetLastError());
    return HQS_OK;
}

int hqs_dag_load(hqs_ctx* ctx, uint32_t n_tasks, const uint32_t* class_id, const uint64_t* priority,
                 const uint32_t* n_deps, const uint32_t* cons_off, const uint32_t* cons) {
    if (!ctx) return HQS_E_INVALID;
    if (!n_tasks || !class_id || !priority || !n_deps || !cons_off) return fail(ctx, HQS_E_INVALID, "null DAG arrays");
    if (ctx->Q == 0) return fail(ctx, HQS_E_STATE, "hqs_classes_set has not been called");
    const u32 n_edges = cons_off[n_tasks];
    if (n_edges && !cons) return fail(ctx, HQS_E_INVALID, "null consumer array");
    for (u32 i = 0; i < n_tasks; ++i)
        if (class_id[i] >= ctx->Q) return fail(ctx, HQS_E_INVALID, "class id %u >= n_classes %u", class_id[i], ctx->Q);
    CU(cudaSetDevice(ctx->device));
    int rc = ensure_handles(ctx, n_tasks);
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->d_deps) { CU(cudaFree(ctx->d_deps)); ctx->d_deps = nullptr; }
    if (ctx->d_cons_off) { CU(cudaFree(ctx->d_cons_off)); ctx->d_cons_off = nullptr; }
    if (ctx->d_cons) { CU(cudaFree(ctx->d_cons)); ctx->d_cons = nullptr; }
    u32* d_cls = nullptr;
    CU(cudaMalloc(&ctx->d_deps, (size_t)n_tasks * 4));
    CU(cudaMalloc(&ctx->d_cons_off, ((size_t)n_tasks + 1) * 4));
    CU(cudaMalloc(&ctx->d_cons, std::max<size_t>(n_edges, 1) * 4));
    CU(cudaMalloc(&d_cls, (size_t)n_tasks * 4));
    CU(cudaMemsetAsync(ctx->d_key, 0, (size_t)ctx->cap_handles * 4, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_deps, n_deps, (size_t)n_tasks * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_cons_off, cons_off, ((size_t)n_tasks + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (n_edges) CU(cudaMemcpyAsync(ctx->d_cons, cons, (size_t)n_edges * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_cls, class_id, (size_t)n_tasks * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_prio, priority, (size_t)n_tasks * 8, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<u64> fresh;
    distinct_priorities(priority, n_tasks, fresh);
    ctx->levels.clear();
    merge_levels(ctx, fresh);
    if ((rc = upload_levels(ctx))) { cudaFree(d_cls); return rc; }
    ctx->n_handles = n_tasks;
    ctx->stats.n_handles = n_tasks;
    dag_init_k<<<(n_tasks + 255) / 256, 256, 0, ctx->stream>>>(n_tasks, d_cls, ctx->d_prio, ctx->d_deps, ctx->d_key,
                                                               ctx->d_levels, (u32)ctx->dev_levels.size(),
                                                               ctx->coarse ? 1 : 0);
    ctx->stats.kernel_launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaFree(d_cls));
    ctx->dag = true;
    return HQS_OK;
}

int hqs_tasks_finished(hqs_ctx* ctx, uint32_t n, const uint32_t* task, uint32_t* n_new_ready) {
    if (!ctx) return HQS_E_INVALID;
    if (n_new_ready) *n_new_ready = 0;
    if (!ctx->dag) return fail(ctx, HQS_E_STATE, "hqs_tasks_finished needs hqs_dag_load");
    if (n == 0) return HQS_OK;
    if (!task) return fail(ctx, HQS_E_INVALID, "null task array");
    CU(cudaSetDevice(ctx->device));
    if (n > ctx->push_cap) {
        CU(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_push_task) { CU(cudaFree(ctx->d_push_task)); CU(cudaFree(ctx->d_push_cls)); CU(cudaFree(ctx->d_push_prio)); }
        ctx->push_cap = std::max<u32>(n, 1u << 16);
        CU(cudaMalloc(&ctx->d_push_task, (size_t)ctx->push_cap * 4));
        CU(cudaMalloc(&ctx->d_push_cls, (size_t)ctx->push_cap * 4));
        CU(cudaMalloc(&ctx->d_push_prio, (size_t)ctx->push_cap * 8));
    }
    CU(cudaMemcpyAsync(ctx->d_push_task, task, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(ctx->d_newcnt, 0, sizeof(u32), ctx->stream));
    finished_k<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, ctx->d_push_task, ctx->d_cons_off, ctx->d_cons, ctx->d_deps,
                                                         ctx->d_key, ctx->d_newcnt);
    ctx->stats.kernel_launches++;
    CU(cudaGetLastError());
    if (n_new_ready) {
        CU(cudaMemcpyAsync(ctx->h_small, ctx->d_newcnt, sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        *n_new_ready = ctx->h_small[0];
    }
    return HQS_OK;
}

int hqs_tick_launch(hqs_ctx* ctx, uint32_t n_workers, const hqs_worker* workers, const uint64_t* free_rw,
                    const uint64_t* total_rw, const uint8_t* blocked_wcv, uint32_t out_cap) {
    if (!ctx) return HQS_E_INVALID;
    int rc = validate_workers(ctx, n_workers, workers, free_rw, total_rw);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    const TickGeom t = tick_geom(ctx);
    if (t.G > HQS_MAX_GROUPS) return fail(ctx, HQS_E_LIMIT, "groups=%u > %u", t.G, HQS_MAX_GROUPS);
    if ((rc = ensure_tick_buffers(ctx, t.G, t.P, n_workers, out_cap))) return rc;
    TickLayout lay;
    if ((rc = upload_tick_input(ctx, n_workers, workers, free_rw, total_rw, blocked_wcv, &lay))) return rc;
    if (ctx->n_handles == 0) {
        // nothing was ever pushed: an empty tick still has to produce a header
        CU(cudaMemsetAsync(ctx->d_hdr, 0, sizeof(TickHeaderOut), ctx->stream));
        CU(cudaMemcpyAsync(ctx->d_free_after, ctx->d_tickin + lay.off_free, (size_t)n_workers * ctx->R * 8,
                           cudaMemcpyDeviceToDevice, ctx->stream));
        ctx->last_W = n_workers; ctx->tick_pending = true; ctx->stats.ticks++;
        return HQS_OK;
    }
    if ((rc = launch_count(ctx, t))) return rc;
    return launch_solve_emit(ctx, t, n_workers, lay, blocked_wcv != nullptr, nullptr, nullptr, out_cap);
}

int hqs_tick_fetch(hqs_ctx* ctx, uint32_t out_cap, hqs_assignment* out, uint32_t* out_n, uint64_t* free_after) {
    if (!ctx) return HQS_E_INVALID;
    if (!ctx->tick_pending) return fail(ctx, HQS_E_STATE, "no tick in flight");
    if (out_n) *out_n = 0;
    CU(cudaSetDevice(ctx->device));
    const size_t fa_bytes = (size_t)ctx->last_W * ctx->R * 8;
    const size_t need = sizeof(TickHeaderOut) + fa_bytes;
    if (need > ctx->h_hdr_cap) {
        if (ctx->h_hdr) CU(cudaFreeHost(ctx->h_hdr));
        ctx->h_hdr_cap = need * 2;
        CU(cudaMallocHost(&ctx->h_hdr, ctx->h_hdr_cap));
    }
    CU(cudaMemcpyAsync(ctx->h_hdr, ctx->d_hdr, sizeof(TickHeaderOut), cudaMemcpyDeviceToHost, ctx->stream));
    if (free_after)
        CU(cudaMemcpyAsync(ctx->h_hdr + sizeof(TickHeaderOut), ctx->d_free_after, fa_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->tick_pending = false;
    TickHeaderOut hdr;
    memcpy(&hdr, ctx->h_hdr, sizeof hdr);
    ctx->stats.n_groups = hdr.n_groups;
    ctx->stats.n_levels = ctx->last_L;
    ctx->stats.n_assigned = hdr.n_assigned;
    ctx->stats.n_segments = hdr.n_segments;
    memcpy(ctx->dbg, hdr.dbg, sizeof ctx->dbg);
    if (hdr.error == 2) return fail(ctx, HQS_E_CUDA, "solver grid synchronisation timed out");
    if (hdr.error) return fail(ctx, HQS_E_LIMIT, "count-segment overflow (> %u segments in one tick)", SEG_CAP);
    if (hdr.n_assigned > out_cap || (hdr.n_assigned && !out))
        return fail(ctx, HQS_E_OVERFLOW, "out_cap=%u too small for %u assignments", out_cap, hdr.n_assigned);
    if (free_after) memcpy(free_after, ctx->h_hdr + sizeof(TickHeaderOut), fa_bytes);
    if (hdr.n_assigned) {
        CU(cudaMemcpyAsync(out, ctx->d_out, (size_t)hdr.n_assigned * sizeof(hqs_assignment), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    if (out_n) *out_n = hdr.n_assigned;
    return HQS_OK;
}

int hqs_tick(hqs_ctx* ctx, uint32_t n_workers, const hqs_worker* workers, const uint64_t* free_rw,
             const uint64_t* total_rw, const uint8_t* blocked_wcv, uint32_t out_cap, hqs_assignment* out,
             uint32_t* out_n, uint64_t* free_after) {
    int rc = hqs_tick_launch(ctx, n_workers, workers, free_rw, total_rw, blocked_wcv, out_cap);
    if (rc) return rc;
    return hqs_tick_fetch(ctx, out_cap, out, out_n, free_after);
}

int hqs_query(hqs_ctx* ctx, uint32_t n_workers, const hqs_worker* workers, const uint64_t* free_rw,
              const uint64_t* total_rw, const uint8_t* blocked_wcv, uint32_t* n_would_assign,
              uint32_t* per_worker_assigned, uint64_t* free_after) {
    if (!ctx) return HQS_E_INVALID;
    if (n_would_assign) *n_would_assign = 0;
    int rc = validate_workers(ctx, n_workers, workers, free_rw, total_rw);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    const TickGeom t = tick_geom(ctx);
    if (t.G > HQS_MAX_GROUPS) return fail(ctx, HQS_E_LIMIT, "groups=%u > %u", t.G, HQS_MAX_GROUPS);
    if ((rc = ensure_tick_buffers(ctx, t.G, t.P, n_workers, 1024))) return rc;
    TickLayout lay;
    if ((rc = upload_tick_input(ctx, n_workers, workers, free_rw, total_rw, blocked_wcv, &lay))) return rc;
    if (ctx->n_handles == 0) {
        if (per_worker_assigned) memset(per_worker_assigned, 0, n_workers * sizeof(u32));
        if (free_after) memcpy(free_after, free_rw, (size_t)n_workers * ctx->R * 8);
        return HQS_OK;
    }
    if ((rc = launch_count(ctx, t))) return rc;
    if ((rc = launch_solve_emit(ctx, t, n_workers, lay, blocked_wcv != nullptr, nullptr, nullptr, 0, true))) return rc;
    ctx->tick_pending = false;      // nothing was emitted or consumed
    ctx->stats.ticks--;
    u32* d_pw = ctx->d_pk_quota;    // scratch (pack is over): [W] counters
    CU(cudaMemsetAsync(d_pw, 0, n_workers * sizeof(u32), ctx->stream));
    seg_worker_totals_k<<<(t.G + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_hdr, ctx->d_gout, t.G, ctx->d_seg_cum, ctx->d_seg_wv, d_pw);
    ctx->stats.kernel_launches++;
    CU(cudaGetLastError());
    TickHeaderOut hdr;
    std::vector<u32> pw(n_workers);
    CU(cudaMemcpyAsync(&hdr, ctx->d_hdr, sizeof hdr, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(pw.data(), d_pw, n_workers * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    if (free_after) CU(cudaMemcpyAsync(free_after, ctx->d_free_after, (size_t)n_workers * ctx->R * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (hdr.error == 2) return fail(ctx, HQS_E_CUDA, "solver grid synchronisation timed out");
    if (hdr.error) return fail(ctx, HQS_E_LIMIT, "count-segment overflow");
    if (n_would_assign) *n_would_assign = hdr.n_assigned;
    if (per_worker_assigned) memcpy(per_worker_assigned, pw.data(), n_workers * sizeof(u32));
    return HQS_OK;
}

int hqs_shard_count(hqs_ctx* ctx, uint32_t n_workers, const hqs_worker* workers, const uint64_t* free_rw,
                    const uint64_t* total_rw, const uint8_t* blocked_wcv, uint32_t* d_counts, uint32_t n_groups_cap,
                    uint32_t* n_groups) {
    if (!ctx) return HQS_E_INVALID;
    int rc = validate_workers(ctx, n_workers, workers, free_rw, total_rw);
    if (rc) return rc;
    if (!d_counts) return fail(ctx, HQS_E_INVALID, "null d_counts");
    CU(cudaSetDevice(ctx->device));
    const TickGeom t = tick_geom(ctx);
    if (t.G > HQS_MAX_GROUPS) return fail(ctx, HQS_E_LIMIT, "groups=%u > %u", t.G, HQS_MAX_GROUPS);
    if (t.G > n_groups_cap) return fail(ctx, HQS_E_LIMIT, "groups=%u > n_groups_cap=%u", t.G, n_groups_cap);
    if ((rc = ensure_tick_buffers(ctx, t.G, t.P, n_workers, 1024))) return rc;
    TickLayout lay;
    if ((rc = upload_tick_input(ctx, n_workers, workers, free_rw, total_rw, blocked_wcv, &lay))) return rc;
    if (ctx->n_handles) {
        if ((rc = launch_count(ctx, t))) return rc;
    }
    CU(cudaMemsetAsync(d_counts, 0, (size_t)n_groups_cap * 4, ctx->stream));
    CU(cudaMemcpyAsync(d_counts, ctx->d_total, (size_t)t.G * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->last_W = n_workers;
    ctx->h_small[8] = blocked_wcv ? 1u : 0u;
    if (n_groups) *n_groups = t.G;
    return HQS_OK;
}

int hqs_shard_solve_emit(hqs_ctx* ctx, const uint32_t* d_counts_all, const uint32_t* d_ranks_before, uint32_t out_cap) {
    if (!ctx) return HQS_E_INVALID;
    if (!d_counts_all || !d_ranks_before) return fail(ctx, HQS_E_INVALID, "null count vectors");
    if (!ctx->last_W) return fail(ctx, HQS_E_STATE, "hqs_shard_count has not been called");
    CU(cudaSetDevice(ctx->device));
    const TickGeom t = tick_geom(ctx);
    int rc = ensure_tick_buffers(ctx, t.G, t.P, ctx->last_W, out_cap);
    if (rc) return rc;
    const TickLayout lay = tick_layout(ctx->last_W, ctx->R, ctx->Q, ctx->h_small[8] != 0);
    return launch_solve_emit(ctx, t, ctx->last_W, lay, ctx->h_small[8] != 0, d_counts_all, d_ranks_before, out_cap);
}

int hqs_tick_reserve(hqs_ctx* ctx, uint32_t n_workers, uint32_t out_cap, int with_blocked) {
    if (!ctx) return HQS_E_INVALID;
    if (n_workers == 0 || n_workers > HQS_MAX_WORKERS) return fail(ctx, HQS_E_LIMIT, "n_workers=%u outside 1..%u", n_workers, HQS_MAX_WORKERS);
    if (ctx->Q == 0) return fail(ctx, HQS_E_STATE, "hqs_classes_set has not been called");
    CU(cudaSetDevice(ctx->device));
    const TickGeom t = tick_geom(ctx);
    if (t.G > HQS_MAX_GROUPS) return fail(ctx, HQS_E_LIMIT, "groups=%u > %u", t.G, HQS_MAX_GROUPS);
    int rc = ensure_tick_buffers(ctx, t.G, t.P, n_workers, out_cap);
    if (rc) return rc;
    const TickLayout lay = tick_layout(n_workers, ctx->R, ctx->Q, with_blocked != 0);
    if (lay.bytes > ctx->tickin_cap) {
        CU(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_tickin) CU(cudaFree(ctx->d_tickin));
        if (ctx->h_tickin) CU(cudaFreeHost(ctx->h_tickin));
        ctx->tickin_cap = lay.bytes * 2;
        CU(cudaMalloc(&ctx->d_tickin, ctx->tickin_cap));
        CU(cudaMallocHost(&ctx->h_tickin, ctx->tickin_cap));
    }
    if (!ctx->h_hdr) {
        ctx->h_hdr_cap = sizeof(TickHeaderOut) + (size_t)HQS_MAX_WORKERS * HQS_MAX_RESOURCES * 8;
        CU(cudaMallocHost(&ctx->h_hdr, ctx->h_hdr_cap));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return HQS_OK;
}

static size_t xbuf_bytes() { return ((size_t)2 * HQS_MAX_PEERS * HQS_MAX_GROUPS + 2 * HQS_MAX_PEERS) * sizeof(u32); }

int hqs_shard_xbuf(hqs_ctx* ctx, void** d_xbuf, uint8_t ipc_handle[HQS_IPC_HANDLE_BYTES]) {
    if (!ctx) return HQS_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    if (!ctx->d_xbuf) {
        CU(cudaMalloc(&ctx->d_xbuf, xbuf_bytes()));
        CU(cudaMemset(ctx->d_xbuf, 0, xbuf_bytes()));
        CU(cudaMalloc(&ctx->d_xall, HQS_MAX_GROUPS * sizeof(u32)));
        CU(cudaMalloc(&ctx->d_xbefore, HQS_MAX_GROUPS * sizeof(u32)));
    }
    if (d_xbuf) *d_xbuf = ctx->d_xbuf;
    if (ipc_handle) {
        static_assert(sizeof(cudaIpcMemHandle_t) == HQS_IPC_HANDLE_BYTES, "IPC handle size");
        cudaIpcMemHandle_t h;
        CU(cudaIpcGetMemHandle(&h, ctx->d_xbuf));
        memcpy(ipc_handle, &h, sizeof h);
    }
    return HQS_OK;
}

int hqs_ipc_open(hqs_ctx* ctx, const uint8_t ipc_handle[HQS_IPC_HANDLE_BYTES], void** d_ptr) {
    if (!ctx || !ipc_handle || !d_ptr) return HQS_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof h);
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->x_opened.push_back(p);
    *d_ptr = p;
    return HQS_OK;
}

int hqs_shard_attach(hqs_ctx* ctx, uint32_t world, uint32_t rank, void* const* peer_xbufs) {
    if (!ctx || !peer_xbufs) return HQS_E_INVALID;
    if (world < 1 || world > HQS_MAX_PEERS || rank >= world) return fail(ctx, HQS_E_LIMIT, "world=%u rank=%u outside 1..%u", world, rank, HQS_MAX_PEERS);
    if (!ctx->d_xbuf) return fail(ctx, HQS_E_STATE, "hqs_shard_xbuf has not been called");
    if (peer_xbufs[rank] != ctx->d_xbuf) return fail(ctx, HQS_E_INVALID, "peer_xbufs[rank] must be this context's own buffer");
    CU(cudaSetDevice(ctx->device));
    for (u32 r = 0; r < world; ++r) {
        if (!peer_xbufs[r]) return fail(ctx, HQS_E_INVALID, "peer %u has no buffer", r);
        ctx->x_peer[r] = static_cast<u32*>(peer_xbufs[r]);
        // a buffer of another device of THIS process needs peer access (IPC mappings were opened with it)
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, peer_xbufs[r]) == cudaSuccess && at.type == cudaMemoryTypeDevice && at.device != ctx->device) {
            const cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return fail(ctx, HQS_E_CUDA, "no peer access from device %d to device %d: %s", ctx->device, at.device, cudaGetErrorString(e));
        }
        cudaGetLastError();
    }
    ctx->x_world = world; ctx->x_rank = rank; ctx->x_seq = 0;
    return HQS_OK;
}

int hqs_shard_tick_launch(hqs_ctx* ctx, uint32_t n_workers, const hqs_worker* workers, const uint64_t* free_rw,
                          const uint64_t* total_rw, const uint8_t* blocked_wcv, uint32_t out_cap) {
    if (!ctx) return HQS_E_INVALID;
    if (!ctx->x_world) return fail(ctx, HQS_E_STATE, "hqs_shard_attach has not been called");
    int rc = validate_workers(ctx, n_workers, workers, free_rw, total_rw);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    const TickGeom t = tick_geom(ctx);
    if (t.G > HQS_MAX_GROUPS) return fail(ctx, HQS_E_LIMIT, "groups=%u > %u", t.G, HQS_MAX_GROUPS);
    if ((rc = ensure_tick_buffers(ctx, t.G, t.P, n_workers, out_cap))) return rc;
    TickLayout lay;
    if ((rc = upload_tick_input(ctx, n_workers, workers, free_rw, total_rw, blocked_wcv, &lay))) return rc;
    if (ctx->n_handles) {
        if ((rc = launch_count(ctx, t))) return rc;
    } else {
        CU(cudaMemsetAsync(ctx->d_total, 0, (size_t)t.G * 4, ctx->stream));
    }
    // every rank advances the sequence number in lockstep (one sharded tick = one exchange)
    ctx->x_seq += 1;
    XchgArgs x;
    for (u32 r = 0; r < HQS_MAX_PEERS; ++r) x.peer[r] = r < ctx->x_world ? ctx->x_peer[r] : nullptr;
    x.world = ctx->x_world; x.rank = ctx->x_rank; x.seq = ctx->x_seq; x.G = t.G;
    xchg_k<<<ctx->x_world, 256, 0, ctx->stream>>>(ctx->d_total, x);
    ctx->stats.kernel_launches++;
    CU(cudaGetLastError());
    ctx->x_tick = true;
    rc = launch_solve_emit(ctx, t, n_workers, lay, blocked_wcv != nullptr, ctx->d_xall, ctx->d_xbefore, out_cap);
    ctx->x_tick = false;
    return rc;
}

int hqs_device_result(hqs_ctx* ctx, const hqs_assignment** d_out, const uint32_t** d_out_n) {
    if (!ctx) return HQS_E_INVALID;
    if (d_out) *d_out = ctx->d_out;
    if (d_out_n) *d_out_n = ctx->d_hdr ? &ctx->d_hdr->n_assigned : nullptr;
    return HQS_OK;
}

void* hqs_stream(hqs_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int hqs_set_stream(hqs_ctx* ctx, void* cuda_stream) {
    if (!ctx) return HQS_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream && ctx->stream) CU(cudaStreamDestroy(ctx->stream));
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    return HQS_OK;
}

int hqs_set_profile(hqs_ctx* ctx, int on) {
    if (!ctx) return HQS_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    if (on && !ctx->ev[0])
        for (int i = 0; i < 4; ++i) CU(cudaEventCreate(&ctx->ev[i]));
    ctx->profile = on != 0;
    ctx->ev_valid = false;
    return HQS_OK;
}

int hqs_get_kernel_ms(hqs_ctx* ctx, float out_ms[4]) {
    if (!ctx || !out_ms) return HQS_E_INVALID;
    if (!ctx->profile || !ctx->ev_valid) return fail(ctx, HQS_E_STATE, "no profiled tick available");
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventSynchronize(ctx->ev[3]));
    out_ms[3] = 0;
    for (int i = 0; i < 3; ++i) {
        CU(cudaEventElapsedTime(&out_ms[i], ctx->ev[i], ctx->ev[i + 1]));
        out_ms[3] += out_ms[i];
    }
    return HQS_OK;
}

int hqs_sync(hqs_ctx* ctx) {
    if (!ctx) return HQS_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    return HQS_OK;
}

int hqs_debug_read(hqs_ctx* ctx, uint64_t out[8]) {
    if (!ctx || !out) return HQS_E_INVALID;
    for (int i = 0; i < 8; ++i) out[i] = ctx->dbg[i];
    return HQS_OK;
}

int hqs_get_stats(hqs_ctx* ctx, hqs_stats* out) {
    if (!ctx || !out) return HQS_E_INVALID;
    *out = ctx->stats;
    out->n_levels = (u32)ctx->dev_levels.size();
    return HQS_OK;
}

}  // extern "C"
