"""Quick device check of the tick kernel: a few small ticks against the specification, then the phase lengths of the
solver CTA (hqs_debug_read) on the benchmark shape.  Usage: python tools/tick_probe.py [n_tasks] [n_workers]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import greedy_model as G
import workloads as WL
from hyperqueue_b200 import _lib as L

if os.environ.get("HQS_LIB"):        # a variant build (tools/variant_build.sh)
    L.LIB_PATH = os.path.join(ROOT, "hyperqueue_b200", os.environ["HQS_LIB"])


def dbg(s):
    d = (C.c_uint64 * 8)()
    s._lib.hqs_debug_read(s._ctx, d)
    return list(d)


def main():
    for (n, w, q, seed, kw) in [(1, 1, 1, 0, {}), (1000, 8, 6, 2, {}), (4097, 16, 12, 3, {}), (20000, 32, 12, 7, dict(variants3=True, blocked_density=0.05)),
                                (60000, 32, 16, 8, dict(free_scale=1024))]:
        wl = WL.make_independent(n, w, q, seed, **kw)
        s = WL.gpu_scheduler(wl)
        fb = s.free.copy()
        t0 = time.perf_counter()
        m = s.run_scheduling()
        dt = time.perf_counter() - t0
        exp, exp_free = G.model_tick(wl, np.ones(wl.n_tasks, dtype=bool), fb)
        ok = np.array_equal(m.assignments, exp) and np.array_equal(m.free_after, exp_free)
        print(f"n={n} w={w} q={q} {kw}: assigned {m.n_assigned()} spec {exp.shape[0]} equal={ok} host {dt*1e3:.2f} ms dbg={dbg(s)}", flush=True)
        if not ok:
            bad = np.nonzero(m.assignments[: min(len(exp), m.n_assigned())] != exp[: min(len(exp), m.n_assigned())])[0]
            print("   first mismatch at", bad[:5], m.assignments[bad[:3]], exp[bad[:3]], flush=True)
        s.close()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    w = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    for scale, tag in [(1024, "M1"), (1, "M2 first tick")]:
        wl = WL.make_independent(n, w, 16, seed=0, free_scale=scale)
        s = WL.gpu_scheduler(wl)
        s._check(s._lib.hqs_set_profile(s._ctx, 1))
        for it in range(4):
            s.free = wl.worker_free.copy()
            t0 = time.perf_counter()
            m = s.run_scheduling()
            dt = time.perf_counter() - t0
            ms = (C.c_float * 4)()
            s._lib.hqs_get_kernel_ms(s._ctx, ms)
            d = dbg(s)
            ghz = d[5] / max(d[6], 1)
            t_gen = ((d[4] >> 16) & 0xFFFF) * 256 / ghz / 1e3
            t_pack = ((d[7] >> 8) & 0xFFFFFF) * 256 / ghz / 1e3
            print(f"{tag} n={n} w={w}: assigned {m.n_assigned()} host {dt*1e3:.3f} ms kernel {ms[3]*1e3:.1f} us "
                  f"[stage+count {d[0]/ghz/1e3:.1f} | compact {d[1]/ghz/1e3:.1f} | solve {d[2]/ghz/1e3:.1f} (pack {t_pack:.1f}, general loop {t_gen:.1f}) | "
                  f"emit {d[3]/ghz/1e3:.1f}] us groups {d[4] & 0xFFFF} visits {d[4] >> 32} packs {d[7] & 0xFF} lean groups {d[7] >> 32} "
                  f"clock {ghz:.3f} GHz", flush=True)
            s.rearm()
        s.close()


if __name__ == "__main__":
    main()
