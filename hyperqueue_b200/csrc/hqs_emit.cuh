// hqs_emit.cuh — emit_k (stable rank -> placement) and the per-worker totals of the what-if query.
// Included by hqsched.cu inside its anonymous namespace.
#pragma once

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
