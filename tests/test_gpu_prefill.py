"""Proactive filling + retract / redirect on the device (hqs_prefill_config / hqs_prefill_state, kind 1 / 2 records) through
GpuScheduler: every tick of the reference's scenarios (tests/prefill_scenarios.py) must equal the specification record by
record, and the per-worker messages must have the reference's shape (prefills first, then assigned; RetractTasks first)."""
import numpy as np
import pytest

import prefill_scenarios as S

pytestmark = pytest.mark.gpu
FR = S.FR


def run_gpu(name):
    from hyperqueue_b200 import GpuScheduler, RequestVariant, priority_from_user
    reserve, pmax, cpus, steps = S.SCENARIOS[name]
    s = GpuScheduler(1)
    c = s.get_or_create_resource_rq_id([RequestVariant.of({0: cpus * FR})])
    s.set_prefill(reserve, pmax)
    n_w = n_t = 0
    recs, maps = [], []
    for new_w, new_t in steps:
        for cp in new_w:
            s.new_worker(50 + n_w, [cp * FR]); n_w += 1
        if new_t:
            h = np.arange(n_t, n_t + new_t, dtype=np.uint32)
            s.add_ready_tasks(h, np.full(new_t, c, dtype=np.uint32), priority_from_user(np.zeros(new_t)))
            n_t += new_t
        m = s.run_scheduling()
        recs.append(m.assignments.copy()); maps.append(m)
    return s, recs, maps


@pytest.mark.parametrize("name", sorted(S.SCENARIOS))
def test_device_equals_specification_with_prefill(name):
    s, recs, maps = run_gpu(name)
    _, exp = S.run_spec(name)
    for tick, (got, want) in enumerate(zip(recs, exp)):
        assert np.array_equal(got, want), (name, tick, got[:8], want[:8])
    s.close()


def test_messages_have_the_reference_shape():
    s, recs, maps = run_gpu("prefill_basic")               # test_scheduler_sn.rs:1168-1200
    msgs = maps[0].messages()
    assert sorted(msgs) == [50, 51]
    for w in (50, 51):
        comp = msgs[w]["compute"]
        assert len(comp) == 34 and all(v is None for _, v in comp[:32]) and all(v == 0 for _, v in comp[32:])
        assert msgs[w]["retracts"] == [] and s.prefilled_tasks(w).size == 32
    s.close()
    s, recs, maps = run_gpu("no_deps_distribute")          # :849-871
    assert sorted(len(m["compute"]) for m in maps[0].messages().values()) == [30, 30, 30]
    s.close()


def test_retract_redirect_and_response():
    """test_prefill_steal (test_scheduler_sn.rs:1225-1306): the second tick takes 2 prefilled tasks away from w1 (RetractTasks)
    and redirects them to w2, which gets 3 fresh tasks right away and the redirected ones when w1 has answered."""
    s, recs, maps = run_gpu("prefill_steal")
    m = maps[1].messages()
    assert len(m[50]["retracts"]) == 2 and m[50]["compute"] == []
    assert len(m[51]["compute"]) == 3 and m[51]["retracts"] == []
    assert sorted(s.redirects.values()) == [(51, 0), (51, 0)]
    assert s.prefilled_tasks(50).size == 3 and s.prefilled_tasks(51).size == 0
    assert int(s.free[1, 0]) == 0                          # w2's 5 cpus are taken: 3 assigned + 2 redirected
    t = sorted(s.redirects)[0]
    sent = s.on_retract_response(50, [t])
    assert sent == {51: [(t, 0)]} and len(s.redirects) == 1
    assert s.on_retract_response(51, [t]) == {}            # not being retracted from that worker: ignored
    s.close()


def test_worker_starts_a_prefilled_task_and_dispose():
    from hyperqueue_b200 import priority_from_user
    s, recs, maps = run_gpu("prefill_choose_waiting")
    w1_pf = s.prefilled_tasks(50)
    assert w1_pf.size == 6
    # the worker cannot start it (its only cpu is busy) in the reference; here only the bookkeeping is exercised: the task
    # leaves the ready set, so the next tick does not hand it out again
    s.free[0, 0] = 1 * FR
    s.on_task_running_prefilled(int(w1_pf[0]), 0)
    assert s.prefilled_tasks(50).size == 5 and int(s.free[0, 0]) == 0
    # a task of higher priority arrives: check_dispose_prefill retracts the class's prefills
    ret = s.dispose_prefill(0)
    assert sorted(len(v) for v in ret.values()) == [4, 5] and s.prefilled_tasks(50).size == 0
    m = s.run_scheduling()
    assert (m.assignments["kind"] != 2).all()              # nothing is prefilled any more, so nothing is redirected
    s.close()
