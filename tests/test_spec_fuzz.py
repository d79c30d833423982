"""CPU-side fuzz of the device algorithm's sequential specification (tests/greedy_model.py): on random ticks that mix
`All` entries, blocked masks, time limits, partly used workers and amounts with remainders, every placement the
specification emits must pass the oracle's feasibility judge, no task may be placed twice, and the free vectors it
reports must equal an exact replay.  (The CUDA path is compared bit for bit with the same specification on the GPU:
tests/test_gpu_edges.py::test_random_ticks_match_specification uses the same generator.)"""
import numpy as np
import pytest

import greedy_model as G
import parity as P
from oracle import judge as J
from test_gpu_edges import _fuzz_workload


@pytest.mark.parametrize("seed", range(40, 100))
def test_specification_output_is_feasible(seed):
    wl = _fuzz_workload(seed)
    fb = wl.worker_free.copy()
    a, fa = G.model_tick(wl, np.ones(wl.n_tasks, dtype=bool), fb)
    assert P.judge_tick(wl, fb, a).ok
    assert np.unique(a["task"]).size == a.size
    amounts, allm, _, _ = wl.class_tables()
    exp = J.replay_free_after(amounts, allm, fb, wl.worker_total, wl.task_class, a["task"], a["worker"], a["variant"])
    assert np.array_equal(exp, fa)
    # priority order of the emission: a task is never emitted before a task of strictly higher priority
    pr = wl.task_user_priority[a["task"]]
    assert (np.diff(pr.astype(np.int64)) <= 0).all()
