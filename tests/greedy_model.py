"""Sequential CPU model of the DEVICE algorithm (priority-ordered first-fit over (level, class) groups).

Test infrastructure: it is neither the oracle (which restates the reference's MILP) nor a product
fallback.  It exists so that (a) the design can be compared with the oracle on the CPU-only box and
(b) the CUDA path can be checked for bit-exact equality with a 60-line specification of itself.
Mirrors hyperqueue_b200/csrc/hqsched.cu: class_order(), solve_body(), emit_k().
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

from parity import FR, MAXV, Workload

AMOUNT_MAX = (1 << 64) - 1
TIME_INF = (1 << 64) - 1

assignment_dtype = np.dtype([("task", "<u4"), ("worker", "<u2"), ("variant", "u1"), ("kind", "u1")])


def class_order(wl: Workload, free: np.ndarray, total: np.ndarray) -> List[int]:
    W, R = free.shape
    S = [0.0] * R
    T = [0.0] * R
    for w in range(W):
        for r in range(R):
            f = int(free[w, r]); t = int(total[w, r])
            S[r] += 1.0 if f == AMOUNT_MAX else f / 10000.0
            T[r] += 1.0 if t == AMOUNT_MAX else t / 10000.0
    scores = []
    for c, vs in enumerate(wl.classes):
        best = 0.0
        for d in vs:
            s = 0.0
            for r in range(R):
                if S[r] < 1e-6:
                    continue
                if r in d.get("all", ()):
                    s += (T[r] / max(W, 1)) / S[r]
                else:
                    s += (int(d["amounts"].get(r, 0)) / 10000.0) / S[r]
            s *= int(np.round(np.float32(d.get("weight", 1.0)) * np.float32(10000))) / 10000.0
            best = max(best, s)
        scores.append((best, c))
    return [c for _, c in sorted(scores, key=lambda sc: -sc[0])]      # stable


def model_tick(wl: Workload, ready: np.ndarray, free: np.ndarray, levels: Optional[np.ndarray] = None,
               remaining_ms: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
    """Returns (assignments in device emission order, free_after)."""
    W, R = free.shape
    total = wl.worker_total
    fr = [[int(x) for x in free[w]] for w in range(W)]
    prio = wl.task_user_priority.astype(np.int64)
    if levels is None:
        levels = np.unique(prio)[::-1]
    if remaining_ms is None:
        remaining_ms = wl.remaining_ms()
    order = class_order(wl, free, total)
    out = []
    ready_idx = np.nonzero(ready)[0]
    key_p = prio[ready_idx]
    key_c = wl.task_class[ready_idx]
    for lvl in levels.tolist():
        in_lvl = ready_idx[key_p == lvl]
        if in_lvl.size == 0:
            continue
        cls_lvl = wl.task_class[in_lvl]
        for c in order:
            tasks = in_lvl[cls_lvl == c]          # ascending handle
            n = int(tasks.size)
            if n == 0:
                continue
            remaining = n
            pos = 0
            for v, d in enumerate(wl.classes[c]):
                if remaining == 0:
                    break
                amounts = {r: int(a) for r, a in d["amounts"].items()}
                alls = tuple(d.get("all", ()))
                min_ms = int(round(d.get("min_time_s", 0.0) * 1000))
                for w in range(W):
                    if remaining == 0:
                        break
                    if wl.blocked is not None and wl.blocked[w, c, v]:
                        continue
                    rt = int(remaining_ms[w])
                    if rt != TIME_INF and min_ms > rt:
                        continue
                    cnt = remaining
                    for r in range(R):
                        if r in alls:
                            q = 1 if (int(total[w, r]) != 0 and fr[w][r] == int(total[w, r])) else 0
                            cnt = min(cnt, q)
                        elif r in amounts and fr[w][r] != AMOUNT_MAX:
                            cnt = min(cnt, fr[w][r] // amounts[r])
                    if cnt <= 0:
                        continue
                    for r in range(R):
                        if r in alls:
                            fr[w][r] = 0
                        elif r in amounts and fr[w][r] != AMOUNT_MAX:
                            fr[w][r] -= cnt * amounts[r]
                    for t in tasks[pos: pos + cnt].tolist():
                        out.append((t, w, v, 0))
                    pos += cnt
                    remaining -= cnt
    a = np.array(out, dtype=assignment_dtype) if out else np.zeros(0, dtype=assignment_dtype)
    return a, np.array(fr, dtype=np.uint64)


def model_drain(wl: Workload, max_ticks: int = 100000):
    """Zero-duration drain with the model (independent tasks or DAG)."""
    n = wl.n_tasks
    if wl.deps is None:
        ready = np.ones(n, dtype=bool)
        unfinished = None
    else:
        unfinished = np.array([len(d) for d in wl.deps], dtype=np.int64)
        ready = unfinished == 0
        consumers = [[] for _ in range(n)]
        for t, ds in enumerate(wl.deps):
            for d in ds:
                consumers[d].append(t)
    levels = np.unique(wl.task_user_priority.astype(np.int64))[::-1]
    remaining = n
    per_tick = []
    while remaining > 0 and len(per_tick) < max_ticks:
        a, _ = model_tick(wl, ready, wl.worker_free, levels)
        if a.size == 0:
            raise RuntimeError("model drain stalled")
        ready[a["task"]] = False
        if unfinished is not None:
            for t in a["task"].tolist():
                for c in consumers[t]:
                    unfinished[c] -= 1
                    if unfinished[c] == 0:
                        ready[c] = True
        remaining -= a.size
        per_tick.append(int(a.size))
    return len(per_tick), per_tick
