"""Workers partly occupied at every tick start (tasks run 1-3 ticks): the device algorithm's specification must not
finish later than 2 % after the oracle (tests/golden/duration_drains.json, generator next to it)."""
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "duration_drains.json")))


@pytest.mark.parametrize("key", sorted(GOLDEN))
def test_specification_keeps_up_with_oracle_under_partial_occupancy(key):
    import make_duration_drains as D
    import parity as P
    g = GOLDEN[key]
    wl = P.make_independent(*g["args"], **g["kwargs"])
    ticks = D.spec_run(wl, D.durations(wl.n_tasks, g["args"][-1]))
    assert ticks == g["model_ticks"]
    assert ticks - g["oracle_ticks"] <= 0.02 * g["oracle_ticks"], (ticks, g)
