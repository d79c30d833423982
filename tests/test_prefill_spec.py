"""Proactive filling and retract / redirect (scheduler/mapping.rs:63-101, 156-230): the specification of the device
algorithm against the restated reference on the reference's own scenarios — per tick and worker the number of assigned
tasks, prefilled tasks and retracted tasks must agree (which tasks are picked inside one class is the same rule in both:
waiting tasks in ascending id first, then prefilled ones)."""
import pytest

import prefill_scenarios as S


@pytest.mark.parametrize("name", sorted(S.SCENARIOS))
def test_specification_matches_oracle_messages(name):
    want = S.run_oracle(name)
    got, _ = S.run_spec(name)
    assert got == want, (got, want)


def test_reference_pinned_numbers():
    got, _ = S.run_spec("prefill_basic")
    assert got == [{0: (2, 32, 0), 1: (2, 32, 0)}]                         # 34 tasks per worker, the first 32 prefills (:1168-1200)
    got, _ = S.run_spec("no_deps_distribute")
    assert got == [{0: (10, 20, 0), 1: (10, 20, 0), 2: (10, 20, 0)}]       # 30 per worker (:849-871)
    got, _ = S.run_spec("prefill_choose_waiting")
    assert [sorted(t.items()) for t in got][-1] == [(0, (0, 0, 0)), (1, (0, 0, 0)), (2, (1, 0, 0))]
    got, _ = S.run_spec("prefill_steal")
    assert got[0] == {0: (1, 5, 0)} and got[1] == {0: (0, 0, 2), 1: (3, 0, 0)}   # 2 retracts, 3 fresh tasks (:1225-1306)
