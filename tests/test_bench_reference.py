"""bench.py's reference arm (the oracle on the host CPU) where mode M1 leaves the range the reference was written for:
a worker that fits more than 1024 tasks of a class reaches the batch limit (workerload.rs:12, batches.rs:80-91), the MILP
then has no size row for the class (solver.rs:245-252) and may hand out more tasks than the class's queue holds.  The
restated reference fails exactly where the real server would panic (take_tasks, taskqueue.rs:326); the bench harness
cuts the counts back to the queue length and keeps the tick alive."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench as B  # noqa: E402
import parity as P  # noqa: E402

CFG = {"tasks_per_gpu": 20000, "workers": 2, "free_scale": 1024, "workload": "unit", "name": "unit"}


def test_restated_reference_overdraws_the_queue_like_the_real_server_would():
    wl = B.make_workload(CFG, 0)
    core = P.oracle_core(wl)
    core.scheduler_state.config.proactive_filling_max = 0
    with pytest.raises(IndexError):            # Rust: Option::unwrap on an empty queue
        core.schedule_mapping(0.0)


def test_bench_harness_cuts_the_counts_back_to_the_queue_length():
    before = getattr(B.oracle_step, "truncated", 0)
    n, seconds, capped = B.oracle_step(CFG, 0)
    assert 0 < n <= CFG["tasks_per_gpu"]
    assert getattr(B.oracle_step, "truncated", 0) == before + 1
    # every assigned task exactly once
    assert seconds > 0.0 and capped in (False, True)
