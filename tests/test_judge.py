"""The feasibility judge (oracle/judge.py) and its plain-C restatement (oracle/judge.c) agree, accept
the oracle's own ticks and reject corrupted ones.  CPU only."""
import ctypes as C
import os

import numpy as np

import parity as P
from oracle import judge as J

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _c_judge(wl, free, a_task, a_worker, a_variant):
    import __graft_entry__ as ge
    ge.build()
    lib = C.CDLL(os.path.join(ROOT, "oracle", "libhqjudge.so"))
    lib.hq_judge.restype = C.c_int64
    amounts, allm, nvar, mint = wl.class_tables()
    W, R = free.shape
    Q, V = amounts.shape[:2]
    blocked = None if wl.blocked is None else np.ascontiguousarray(wl.blocked.astype(np.uint8))
    arrs = [np.ascontiguousarray(amounts), np.ascontiguousarray(allm.astype(np.uint8)), np.ascontiguousarray(mint),
            np.ascontiguousarray(free), np.ascontiguousarray(wl.worker_total), np.ascontiguousarray(wl.remaining_ms())]
    t = np.ascontiguousarray(a_task, dtype=np.uint32); w = np.ascontiguousarray(a_worker, dtype=np.uint16)
    v = np.ascontiguousarray(a_variant, dtype=np.uint8); tc = np.ascontiguousarray(wl.task_class, dtype=np.uint32)
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    return lib.hq_judge(C.c_uint32(W), C.c_uint32(R), C.c_uint32(Q), C.c_uint32(V), *[p(x) for x in arrs],
                        p(blocked) if blocked is not None else None, p(tc), C.c_uint64(t.size), p(t), p(w), p(v))


def test_judge_accepts_oracle_tick_and_rejects_corruption():
    wl = P.make_independent(1500, 6, 5, seed=11, blocked_density=0.1)
    core = P.oracle_core(wl)
    core.scheduler_state.config.proactive_filling_max = 0
    ts, ws, vs, _ = P.oracle_tick(core)
    assert ts.size > 50
    amounts, allm, nvar, mint = wl.class_tables()
    args = (amounts, allm, nvar, mint, wl.worker_free, wl.worker_total, wl.remaining_ms(), wl.blocked, wl.task_class)
    ok = J.judge_assignments(*args, ts, ws, vs)
    assert ok.ok, ok
    assert _c_judge(wl, wl.worker_free, ts, ws, vs) == 0
    # exact replay equals the oracle's tracked free vectors (Worker::sanity_check)
    exp = J.replay_free_after(amounts, allm, wl.worker_free, wl.worker_total, wl.task_class, ts, ws, vs)
    got = np.array([[core.workers[w].free.get(r) for r in range(wl.R)] for w in range(wl.n_workers)], dtype=np.uint64)
    assert np.array_equal(exp, got)
    # over-commit one worker: everything onto worker 0
    bad = J.judge_assignments(*args, ts, np.zeros_like(ws), vs)
    assert not bad.ok
    assert _c_judge(wl, wl.worker_free, ts, np.zeros_like(ws), vs) > 0
    # duplicate task
    dup = J.judge_assignments(*args, np.concatenate([ts, ts[:1]]), np.concatenate([ws, ws[:1]]), np.concatenate([vs, vs[:1]]))
    assert not dup.ok
    # blocked pair
    w0, c0 = int(ws[0]), int(wl.task_class[ts[0]])
    wl.blocked[w0, c0, 0] = True
    blk = J.judge_assignments(amounts, allm, nvar, mint, wl.worker_free, wl.worker_total, wl.remaining_ms(), wl.blocked,
                              wl.task_class, ts, ws, vs)
    assert not blk.ok and _c_judge(wl, wl.worker_free, ts, ws, vs) > 0


def test_judge_all_policy_and_time():
    FR = P.FR
    classes = [[{"amounts": {}, "all": (0,)}], [{"amounts": {0: 2 * FR}, "min_time_s": 50.0}]]
    total = np.array([[4 * FR], [4 * FR]], dtype=np.uint64)
    free = np.array([[4 * FR], [3 * FR]], dtype=np.uint64)
    wl = P.Workload(1, classes, total, free, np.array([0, 0, 1, 1], dtype=np.uint32), np.zeros(4, dtype=np.int32),
                    worker_remaining_s=np.array([np.inf, 10.0]))
    am, allm, nvar, mint = wl.class_tables()
    base = (am, allm, nvar, mint, free, total, wl.remaining_ms(), None, wl.task_class)
    t = lambda *x: np.array(x, dtype=np.int64)
    assert J.judge_assignments(*base, t(0), t(0), t(0)).ok                 # All on an untouched worker
    assert not J.judge_assignments(*base, t(0), t(1), t(0)).ok             # All consumes the total: 4 > 3 free
    assert not J.judge_assignments(*base, t(0, 1), t(0, 0), t(0, 0)).ok    # two All tasks on one worker
    assert J.judge_assignments(*base, t(2, 3), t(0, 0), t(0, 0)).ok
    assert not J.judge_assignments(*base, t(2), t(1), t(0)).ok             # needs 50 s, 10 s left
