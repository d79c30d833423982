"""Known-answer vectors of the reference's scheduler tests in a neutral form, for running the SAME cases
through the sequential specification (CPU) and the CUDA path (GPU).

Each case: workers (cpus[, extra resources]), tasks (priority, request variants), expected per-worker result as
a multiset of (class key, variant) counts — or None where only a predicate is pinned.  Source lines refer to
/root/reference/crates/tako/src/internal/tests/test_scheduler_sn.rs.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np

from workloads import FR, Workload


def case(workers, tasks, expect, name, running=None, resources=1):
    return {"workers": workers, "tasks": tasks, "expect": expect, "name": name, "running": running or {}, "R": resources}


def c(cpus, **extra):          # single-variant request: cpus + {rid: units}
    d = {0: cpus}
    d.update({int(k[1:]): v for k, v in extra.items()})
    return (tuple(sorted(d.items())),)


def cv(*variants):             # variants: dicts rid -> units
    return tuple(tuple(sorted(v.items())) for v in variants)


CASES = [
    # test_schedule_no_priorities :156-224  (expected: per worker list of cpu sizes)
    case([(3,)], [(0, c(3))], [[c(3)]], "nop-1"),
    case([(4,), (4,)], [(0, c(2))], [[c(2)], []], "nop-2"),
    case([(4,), (4,)], [(0, c(2))] * 2, [[c(2)] * 2, []], "nop-3"),
    case([(4,), (4,)], [(0, c(2))] * 3, [[c(2)] * 2, [c(2)]], "nop-4"),
    case([(4,), (4,)], [(0, c(2))] * 4, [[c(2)] * 2, [c(2)] * 2], "nop-5"),
    case([(4,), (4,)], [(0, c(2))] * 5, [[c(2)] * 2, [c(2)] * 2], "nop-6"),
    case([(4,), (4,)], [(0, c(2)), (0, c(3))], [[c(3)], [c(2)]], "nop-7"),
    case([(3,), (4,)], [(0, c(2)), (0, c(3))], [[c(3)], [c(2)]], "nop-8"),
    case([(4,), (4,)], [(0, c(5))] * 2 + [(0, c(1))] * 5, [[c(1)] * 4, [c(1)]], "nop-9"),
    case([(4,), (4,)], [(0, c(3)), (0, c(4)), (0, c(2))], [[c(4)], [c(3)]], "nop-10"),
    # test_schedule_priorities :226-307
    case([(4,), (4,)], [(1, c(2)), (1, c(2))], [[c(2)] * 2, []], "prio-1"),
    case([(4,), (4,)], [(1, c(2)), (2, c(2))], [[c(2)] * 2, []], "prio-2"),
    case([(4,), (4,)], [(0, c(4)), (0, c(4)), (1, c(2)), (2, c(3))], [[c(3)], [c(2)]], "prio-3"),
    case([(4,), (4,)], [(0, c(4)), (0, c(4)), (1, c(2)), (1, c(3))], [[c(3)], [c(2)]], "prio-4"),
    case([(4,), (4,)], [(1, c(4)), (1, c(4)), (1, c(2)), (1, c(3))], [[c(4)], [c(4)]], "prio-5"),
    case([(4,), (4,)], [(0, c(2)), (4, c(2)), (3, c(1)), (2, c(3))], "set:" , "prio-6-all-four"),     # MILP places all 4
    case([(4,), (4,)], [(1, c(5)), (0, c(4))], [[c(4)], []], "prio-7"),
    case([(4,)], [(9, c(2)), (7, c(1)), (6, c(2))], [[c(2), c(1)]], "prio-9"),
    case([(4,)], [(9, c(2)), (7, c(1)), (6, c(2)), (5, c(1))], [[c(2), c(1)]], "prio-10"),
    case([(10,)], [(9, c(2)), (8, c(1)), (7, c(2)), (6, c(1)), (5, c(2)), (4, c(1)), (3, c(2)), (2, c(1))],
         [[c(2), c(1), c(2), c(1), c(2), c(1)]], "prio-11"),
    case([(4,)], [(1, c(3))] * 3 + [(0, c(1))], [[c(3), c(1)]], "prio-12"),
    # test_schedule_no_irrelevant_blocking :309-330
    case([(3,)], [(10, c(5)), (0, c(1))], [[c(1)]], "noblock-1"),
    case([(3,), (5,)], [(10, c(5)), (9, c(5)), (0, c(1))], [[c(1)], [c(5)]], "noblock-2"),
    case([(5,), (3,)], [(10, c(3)), (9, c(2)), (8, c(5)), (0, c(1))], [[c(3), c(2)], [c(1)]], "noblock-3"),
    # test_schedule_gap_filling :410-449
    case([(12,)], [(1, c(8)), (1, c(8)), (0, c(4))], [[c(8), c(4)]], "gap-1"),
    case([(6,)], [(1, c(3))] * 3 + [(0, c(2))], [[c(3), c(3)]], "gap-2"),
    case([(8,)], [(1, c(3))] * 3 + [(0, c(1))] * 2, [[c(3), c(3), c(1), c(1)]], "gap-3"),
    case([(8,)], [(1, c(3))] * 3 + [(2, c(1)), (0, c(1))], [[c(1), c(3), c(3), c(1)]], "gap-4"),
    # test_schedule_multiple_resources1/2 :635-721 (gpus = resource 1)
    case([(4, 2)], [(0, c(2, r1=1))] * 2, [[c(2, r1=1)] * 2], "mres-1", resources=2),
    case([(4, 1)], [(0, c(2, r1=1))] * 2, [[c(2, r1=1)]], "mres-2", resources=2),
    case([(4, 1)], [(0, c(1, r1=2))], [[]], "mres-5", resources=2),
    case([(6, 0)], [(0, c(2))] * 10 + [(0, c(2, r1=1))] * 10, [[c(2)] * 3], "mres2-1", resources=2),
    case([(6, 10)], [(0, c(2))] * 10 + [(0, c(2, r1=1))] * 10, [[c(2, r1=1)] * 3], "mres2-2", resources=2),
    case([(6, 2)], [(0, c(2))] * 10 + [(0, c(2, r1=1))] * 10, [[c(2, r1=1)] * 2 + [c(2)]], "mres2-3", resources=2),
    case([(6, 2), (6, 0)], [(0, c(2))] * 10 + [(0, c(2, r1=1))] * 10, [[c(2, r1=1)] * 2 + [c(2)], [c(2)] * 3], "mres2-4", resources=2),
    # test_schedule_variants1 :723-754   (variant 0 = 2 cpus, variant 1 = 5 cpus)
    case([(11,)], [(0, cv({0: 2}, {0: 5}))] * 2, "count:2", "var1-1"),
    case([(8,)], [(0, cv({0: 2}, {0: 5}))] * 10, "count:4", "var1-4"),
    # test_generic_resource_variants1-3 :1053-1108
    case([(4, 0), (4, 2)], [(0, cv({0: 2}, {0: 1, 1: 1}))] * 4, "perworker:2,2", "gvar-1", resources=2),
    case([(4, 0), (4, 2)], [(0, cv({0: 8}, {0: 1, 1: 1}))] * 4, "perworker:0,2", "gvar-2", resources=2),
    case([(2, 0), (5, 1)], [(0, cv({0: 3}, {0: 1, 1: 1}))] * 4, "perworker:0,2", "gvar-3", resources=2),
    # test_no_deps_scattering_1 :793-815 (compaction)
    case([(5,), (5,), (5,)], [(0, c(1))] * 4, "perworker:4,0,0", "scatter-1"),
    # test_schedule_running :1308-1322
    case([(14,)], [(0, c(1))] * 10, "count:6", "running-1", running={0: 8}),
    # test_schedule_some_tasks_running :332-366
    case([(3,)], [(1, c(3))], "count:0", "somerun-1", running={0: 1}),
    case([(3,)], [(1, c(2))], "count:1", "somerun-2", running={0: 1}),
]


def to_workload(cs) -> Tuple[Workload, List[Tuple]]:
    R = cs["R"]
    classes: List = []
    keys: List[Tuple] = []
    cls_of = []
    for prio, key in cs["tasks"]:
        if key not in keys:
            keys.append(key)
            classes.append([{"amounts": {r: int(u * FR) for r, u in var}} for var in key])
        cls_of.append(keys.index(key))
    W = len(cs["workers"])
    total = np.zeros((W, R), dtype=np.uint64)
    for w, res in enumerate(cs["workers"]):
        for r, u in enumerate(res):
            total[w, r] = u * FR
    free = total.copy()
    for w, used_cpus in cs["running"].items():
        free[w, 0] -= np.uint64(used_cpus * FR)
    wl = Workload(R, classes, total, free, np.array(cls_of, dtype=np.uint32),
                  np.array([p for p, _ in cs["tasks"]], dtype=np.int32), name=cs["name"])
    return wl, keys


def check(cs, wl, keys, a) -> Optional[str]:
    """a: assignment array (task, worker, variant).  Returns None if the expectation holds, else a message."""
    W = wl.n_workers
    exp = cs["expect"]
    per = [[] for _ in range(W)]
    for t, w, v in zip(a["task"].tolist(), a["worker"].tolist(), a["variant"].tolist()):
        per[w].append(keys[wl.task_class[t]])
    if isinstance(exp, str):
        kind, _, arg = exp.partition(":")
        if kind == "count":
            return None if a.shape[0] == int(arg) else f"assigned {a.shape[0]} != {arg}"
        if kind == "perworker":
            want = [int(x) for x in arg.split(",")]
            got = [len(p) for p in per]
            return None if got == want else f"per-worker {got} != {want}"
        if kind == "set":
            return None if a.shape[0] == len(cs["tasks"]) else f"assigned {a.shape[0]} of {len(cs['tasks'])}"
        raise ValueError(exp)
    got = [sorted(map(repr, p)) for p in per]
    want = [sorted(map(repr, e)) for e in exp]
    return None if got == want else f"{got} != {want}"
