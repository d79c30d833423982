"""Full-size runs of the BASELINE.json configurations on the GPU (no oracle at these sizes; every Nth tick goes
through the feasibility judge).  python tools/config_probe.py [cfg2|cfg3|cfg4|all]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity as P


def drain(wl, judge_every=50, label=""):
    s = P.gpu_scheduler(wl)
    t0 = time.perf_counter()
    left, ticks, t_tick = wl.n_tasks, 0, 0.0
    while left > 0:
        fb = s.free.copy()
        t1 = time.perf_counter()
        m = s.run_scheduling()
        t_tick += time.perf_counter() - t1
        if m.n_assigned() == 0:
            raise RuntimeError("stalled")
        if ticks % judge_every == 0:
            assert P.judge_tick(wl, fb, m.assignments).ok
        left -= m.n_assigned()
        ticks += 1
        s.tasks_finished(m.assignments["task"], propagate=wl.deps is not None)
    dt = time.perf_counter() - t0
    print(f"{label}: {wl.name}: ticks={ticks} total={dt:.2f}s tick-calls={t_tick:.2f}s "
          f"-> {wl.n_tasks / t_tick / 1e6:.2f} M assignments/s in hqs_tick, {1e3 * t_tick / ticks:.3f} ms/tick", flush=True)
    s.close()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("cfg2", "all"):
        drain(P.make_independent(1_000_000, 256, 16, seed=0), label="cfg2 M2 drain")
    if which in ("cfg3", "all"):
        drain(P.make_independent(1_000_000, 256, 16, seed=0, variants3=True, blocked_density=0.05), label="cfg3 M2 drain (3 variants, 5% blocked)")
    if which in ("cfg4", "all"):
        t0 = time.time()
        wl = P.make_dag(500_000, 256, 16, seed=0)
        print("dag built in %.0fs, max b-level %d" % (time.time() - t0, wl.task_user_priority.max()), flush=True)
        drain(wl, label="cfg4 DAG waves")
