"""The reference's own known-answer vectors (tests/vectors.py, transcribed from test_scheduler_sn.rs) run
through the device algorithm: its sequential specification on CPU, the CUDA path on the GPU.

41 of the 44 single-tick vectors are reproduced exactly.  The 3 documented deviations:
  prio-6-all-four  the MILP finds the one arrangement that places all four tasks, first-fit places three
  prio-10, prio-11 the reference's priority cut keeps a lower-priority 1-cpu task out of the gap a waiting
                   2-cpu class could use (gap.rs); gap/reservation semantics are not implemented on the device
"""
import numpy as np
import pytest

import greedy_model as G
import vectors as V

KNOWN_DEVIATIONS = {"prio-6-all-four", "prio-10", "prio-11"}


@pytest.mark.parametrize("cs", V.CASES, ids=[c["name"] for c in V.CASES])
def test_specification_on_reference_vectors(cs):
    wl, keys = V.to_workload(cs)
    a, _ = G.model_tick(wl, np.ones(wl.n_tasks, dtype=bool), wl.worker_free)
    msg = V.check(cs, wl, keys, a)
    if cs["name"] in KNOWN_DEVIATIONS:
        assert msg is not None, "a documented deviation disappeared: update KNOWN_DEVIATIONS and DESIGN.md"
    else:
        assert msg is None, msg


@pytest.mark.gpu
@pytest.mark.parametrize("cs", V.CASES, ids=[c["name"] for c in V.CASES])
def test_cuda_path_on_reference_vectors(cs):
    import parity as P
    wl, keys = V.to_workload(cs)
    s = P.gpu_scheduler(wl)
    fb = s.free.copy()
    m = s.run_scheduling()
    s.close()
    assert P.judge_tick(wl, fb, m.assignments).ok
    msg = V.check(cs, wl, keys, m.assignments)
    if cs["name"] in KNOWN_DEVIATIONS:
        assert msg is not None
    else:
        assert msg is None, msg
