// TEST DOUBLE of the C ABI (include/hqsched.h) — test infrastructure, never shipped or loaded by the product.
// It lets the C++ host shim (hyperqueue_b200/csrc/tako_shim.cpp) be exercised on a box without a GPU: ready set,
// classes and ticks are kept in host memory and a tick is a plain priority-ordered first-fit over variant order.
// Only the entry points the shim calls are implemented.
#include "../../include/hqsched.h"

#include <algorithm>
#include <cstring>
#include <map>
#include <string>
#include <vector>

struct hqs_ctx {
    uint32_t R = 0;
    std::vector<hqs_class> classes;
    struct T { uint32_t cls; uint64_t prio; bool ready; };
    std::map<uint32_t, T> tasks;                 // by handle
    hqs_stats stats{};
    std::string err;
    uint32_t pushes = 0, removes = 0;
};

extern "C" {
int hqs_abi_version(void) { return HQS_ABI_VERSION; }
const char* hqs_last_error(const hqs_ctx* ctx) { return ctx ? ctx->err.c_str() : "fake"; }
int hqs_create(hqs_ctx** out, int, uint32_t n_resources, uint32_t) {
    *out = new hqs_ctx();
    (*out)->R = n_resources;
    return HQS_OK;
}
void hqs_destroy(hqs_ctx* ctx) { delete ctx; }
int hqs_classes_set(hqs_ctx* ctx, uint32_t n, const hqs_class* classes) {
    ctx->classes.assign(classes, classes + n);
    return HQS_OK;
}
int hqs_ready_push(hqs_ctx* ctx, uint32_t n, const uint32_t* task, const uint32_t* class_id, const uint64_t* priority) {
    for (uint32_t i = 0; i < n; ++i) {
        if (class_id[i] >= ctx->classes.size()) { ctx->err = "class id out of range"; return HQS_E_INVALID; }
        ctx->tasks[task[i]] = {class_id[i], priority[i], true};
    }
    ctx->pushes += n;
    return HQS_OK;
}
int hqs_ready_remove(hqs_ctx* ctx, uint32_t n, const uint32_t* task) {
    for (uint32_t i = 0; i < n; ++i) {
        auto it = ctx->tasks.find(task[i]);
        if (it != ctx->tasks.end()) it->second.ready = false;
    }
    ctx->removes += n;
    return HQS_OK;
}
int hqs_get_stats(hqs_ctx* ctx, hqs_stats* out) {
    ctx->stats.n_handles = ctx->tasks.empty() ? 0 : ctx->tasks.rbegin()->first + 1;
    ctx->stats.kernel_launches = ctx->pushes;          // the test reads the push counter through this field
    ctx->stats.n_segments = ctx->removes;
    *out = ctx->stats;
    return HQS_OK;
}
int hqs_tick(hqs_ctx* ctx, uint32_t W, const hqs_worker* workers, const uint64_t* free_rw, const uint64_t* total_rw,
             const uint8_t* blocked, uint32_t out_cap, hqs_assignment* out, uint32_t* out_n, uint64_t* free_after) {
    const uint32_t R = ctx->R, Q = (uint32_t)ctx->classes.size();
    for (uint32_t w = 1; w < W; ++w)
        if (workers[w].worker_id <= workers[w - 1].worker_id) { ctx->err = "workers must be sorted"; return HQS_E_INVALID; }
    std::vector<uint64_t> fr(free_rw, free_rw + (size_t)W * R);
    std::vector<std::pair<uint64_t, uint32_t>> order;   // (priority desc, handle asc)
    for (auto& kv : ctx->tasks)
        if (kv.second.ready) order.push_back({~kv.second.prio, kv.first});
    std::sort(order.begin(), order.end());
    uint32_t n = 0;
    for (auto& po : order) {
        hqs_ctx::T& t = ctx->tasks[po.second];
        const hqs_class& c = ctx->classes[t.cls];
        bool placed = false;
        for (uint32_t w = 0; w < W && !placed; ++w)
            for (uint32_t v = 0; v < c.n_variants && !placed; ++v) {
                const hqs_variant& hv = c.variants[v];
                if (blocked && ((blocked[(size_t)w * Q + t.cls] >> v) & 1)) continue;
                if (workers[w].remaining_time_ms != HQS_TIME_INF && hv.min_time_ms > workers[w].remaining_time_ms) continue;
                bool ok = true;
                for (uint32_t r = 0; r < R; ++r) {
                    if ((hv.all_mask >> r) & 1) ok &= total_rw[(size_t)w * R + r] != 0 && fr[(size_t)w * R + r] == total_rw[(size_t)w * R + r];
                    else ok &= hv.amount[r] <= fr[(size_t)w * R + r];
                }
                if (!ok || n >= out_cap) continue;
                for (uint32_t r = 0; r < R; ++r) {
                    if ((hv.all_mask >> r) & 1) fr[(size_t)w * R + r] = 0;
                    else fr[(size_t)w * R + r] -= hv.amount[r];
                }
                out[n].task = po.second; out[n].worker = (uint16_t)w; out[n].variant = (uint8_t)v; out[n].kind = 0;
                ++n;
                t.ready = false;
                placed = true;
            }
    }
    *out_n = n;
    if (free_after) std::copy(fr.begin(), fr.end(), free_after);
    ctx->stats.n_assigned = n;
    ctx->stats.ticks++;
    return HQS_OK;
}
// proactive filling is not modelled by the double: configuration is accepted, nothing is ever prefilled
int hqs_prefill_config(hqs_ctx*, uint32_t, uint32_t) { return HQS_OK; }
int hqs_prefill_state(hqs_ctx*, uint32_t, const uint8_t*) { return HQS_OK; }
int hqs_prefill_dispose(hqs_ctx*, uint32_t) { return HQS_OK; }
}
