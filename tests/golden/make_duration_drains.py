"""Generates tests/golden/duration_drains.json: makespans (ticks until the last task has finished) of drains in
which every task runs for 1-3 ticks, so that workers are partly occupied (free != total) at every tick start —
the ORACLE (restated reference tick, parity.ORACLE_FAST) next to the device algorithm's sequential specification.

Run from the repo root:  python tests/golden/make_duration_drains.py
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import numpy as np  # noqa: E402

import greedy_model as G  # noqa: E402
import parity as P  # noqa: E402

CASES = {
    "dur_indep_3000_8_6_51": ([3000, 8, 6, 51], {}),
    "dur_indep3_3000_8_6_53": ([3000, 8, 6, 53], {"variants3": True}),
    "dur_indep_6000_12_10_55": ([6000, 12, 10, 55], {}),
}


def durations(n: int, seed: int) -> np.ndarray:
    return np.random.default_rng(seed).integers(1, 4, n)


def spec_run(wl, dur) -> int:
    n = wl.n_tasks
    ready = np.ones(n, dtype=bool)
    free = wl.worker_free.copy()
    amounts, _, _, _ = wl.class_tables()
    levels = np.unique(wl.task_user_priority.astype(np.int64))[::-1]
    remaining, tick, finish_at = n, 0, {}
    while remaining > 0 and tick < 100000:
        for (t, w, v) in finish_at.pop(tick, []):
            free[w] += amounts[wl.task_class[t], v]            # workerload.rs:194-202 (no `All` in these workloads)
        a, fa = G.model_tick(wl, ready, free, levels)
        free = fa.copy()
        ready[a["task"]] = False
        for t, w, v in zip(a["task"].tolist(), a["worker"].tolist(), a["variant"].tolist()):
            finish_at.setdefault(tick + int(dur[t]), []).append((t, w, v))
        remaining -= a.size
        tick += 1
    return tick + max((k - tick for k in finish_at), default=0)


def oracle_run(wl, dur) -> int:
    core = P.oracle_core(wl)
    core.scheduler_state.config.proactive_filling_max = 0
    remaining, tick, finish_at = wl.n_tasks, 0, {}
    while remaining > 0 and tick < 100000:
        for (t, w) in finish_at.pop(tick, []):
            core.task_finished(w, t)
        ts, ws, _, _ = P.oracle_tick(core)
        for t, w in zip(ts.tolist(), ws.tolist()):
            finish_at.setdefault(tick + int(dur[t]), []).append((t, w))
        remaining -= ts.size
        tick += 1
    return tick + max((k - tick for k in finish_at), default=0)


if __name__ == "__main__":
    out = {}
    for key, (args, kw) in CASES.items():
        wl = P.make_independent(*args, **kw)
        dur = durations(wl.n_tasks, args[-1])
        t0 = time.time()
        out[key] = {"args": args, "kwargs": kw, "model_ticks": spec_run(wl, dur), "oracle_ticks": oracle_run(wl, dur),
                    "seconds": round(time.time() - t0, 1)}
        print(key, out[key], flush=True)
        json.dump(out, open(os.path.join(HERE, "duration_drains.json"), "w"), indent=1, sort_keys=True)
