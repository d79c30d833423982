"""BASELINE.json configs[1], [2], [3] at FULL size, mode M2: zero-duration drains through the public call
on the GPU.  The Python specification and the oracle are far too slow at this size (minutes to hours), so the checks are the
size-independent ones: every k-th tick through the reference's feasibility judge + the exact replay of the free vectors,
every task assigned exactly once, resources back to the initial vectors, and the makespan pinned to the value recorded when
the algorithm was validated against the oracle on the scaled fixtures (a regression of the packing rules moves it)."""
import numpy as np
import pytest

import parity as P
from oracle import judge as J

pytestmark = pytest.mark.gpu


def _drain(wl, judge_every, dag=False):
    s = P.gpu_scheduler(wl)
    amounts, allm, _, _ = wl.class_tables()
    seen = np.zeros(wl.n_tasks, dtype=bool)
    ready = np.ones(wl.n_tasks, dtype=bool) if wl.deps is None else np.array([len(d) == 0 for d in wl.deps])
    left, ticks = wl.n_tasks, 0
    while left > 0 and ticks < 20000:
        fb = s.free.copy()
        m = s.run_scheduling()
        a = m.assignments
        assert a.size, f"stalled with {left} tasks left"
        assert not seen[a["task"]].any()
        seen[a["task"]] = True
        if ticks % judge_every == 0:
            assert P.judge_tick(wl, fb, a, ready if wl.deps is None else None).ok
            exp = J.replay_free_after(amounts, allm, fb, wl.worker_total, wl.task_class, a["task"], a["worker"], a["variant"])
            assert np.array_equal(exp, m.free_after)
        ready[a["task"]] = False
        s.tasks_finished(a["task"], propagate=dag)
        left -= a.size
        ticks += 1
    assert left == 0 and seen.all() and np.array_equal(s.free, wl.worker_free)
    s.close()
    return ticks


def test_cfg2_full_size_drain():
    assert _drain(P.make_independent(1_000_000, 256, 16, seed=0), 50) == 491


def test_cfg3_full_size_drain_variants_and_blocked():
    assert _drain(P.make_independent(1_000_000, 256, 16, seed=0, variants3=True, blocked_density=0.05), 50) == 331


def test_cfg4_full_size_dag_drain():
    wl = P.make_dag(500_000, 256, 16, seed=0)
    assert _drain(wl, 100, dag=True) == 1234          # = the critical path of the DAG: one tick per completion wave
