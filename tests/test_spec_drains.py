"""CPU twin of test_gpu_parity.py::test_drain_makespan_vs_oracle: the sequential specification of the device
algorithm (tests/greedy_model.py, which the CUDA path equals bit for bit) must drain every golden workload within
the north_star's 2 % of the oracle's makespan (tests/golden/oracle_drains.json).  Runs without a GPU, so a change of
the algorithm is checked against the oracle before it is ported to the kernels."""
import json
import os

import pytest

import greedy_model as G
import workloads as WL

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_drains.json")
KEYS = sorted(k for k in json.load(open(GOLDEN)).keys() if not k.startswith(("big_", "w256_")) or os.environ.get("HQS_BIG_DRAINS"))


@pytest.mark.parametrize("key", KEYS)
def test_specification_drain_vs_oracle(key):
    golden = json.load(open(GOLDEN))[key]
    wl = (WL.make_dag if "dag" in key.split("_")[:2] else WL.make_independent)(*golden["args"], **golden.get("kwargs", {}))
    ticks, per_tick = G.model_drain(wl)
    assert sum(per_tick) == wl.n_tasks
    assert ticks <= golden["max_ticks"], (ticks, golden["max_ticks"])
    assert ticks - golden["oracle_ticks"] <= 0.02 * golden["oracle_ticks"], (ticks, golden["oracle_ticks"])
