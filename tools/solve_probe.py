"""Times the three tick kernels (CUDA events inside the library) for a few workload shapes."""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import parity as P
from hyperqueue_b200 import _lib as L

def probe(n, w, q, npri, free_scale, reps=5, label=""):
    wl = P.make_independent(n, w, q, seed=0, free_scale=free_scale, n_priorities=npri)
    s = P.gpu_scheduler(wl)
    lib = s._lib
    s._check(lib.hqs_set_profile(s._ctx, 1))
    workers = s._worker_structs(0.0)
    free = np.ascontiguousarray(wl.worker_free); total = np.ascontiguousarray(wl.worker_total)
    out = np.zeros(n, dtype=L.assignment_dtype); out_n = C.c_uint32(0)
    acc = np.zeros(4)
    for i in range(reps + 1):
        s._check(lib.hqs_tick_launch(s._ctx, w, L.ptr(workers), L.ptr(free), L.ptr(total), None, n))
        s._check(lib.hqs_sync(s._ctx))
        ms = (C.c_float * 4)()
        s._check(lib.hqs_get_kernel_ms(s._ctx, ms))
        s._check(lib.hqs_tick_fetch(s._ctx, n, L.ptr(out), C.byref(out_n), None))
        s._check(lib.hqs_ready_rearm(s._ctx))
        if i:
            acc += np.array(list(ms))
    acc /= reps
    dbg = (C.c_uint64 * 8)(); lib.hqs_debug_read(s._ctx, dbg)
    print('   cycles: compact=%d sat=%d groups=%d total=%d nlist=%d' % tuple(dbg[:5]))
    st = s.stats()
    print(f"{label:28s} n={n} w={w} q={q} L={npri} groups={st['n_groups']} segs={st['n_segments']} assigned={out_n.value} "
          f"count={acc[0]*1e3:.1f}us solve={acc[1]*1e3:.1f}us emit={acc[2]*1e3:.1f}us", flush=True)
    s.close()

if __name__ == "__main__":
    probe(1_000_000, 256, 16, 8, 1024, label="M1 headline")
    probe(1_000_000, 256, 16, 1, 1024, label="M1 one level")
    probe(1_000_000, 256, 1, 1, 1024, label="M1 one group")
    probe(1_000_000, 256, 16, 8, 1, label="M2 first tick")
    probe(100_000, 256, 16, 8, 1024, label="M1 100k")
    probe(1_000_000, 32, 16, 8, 1024 * 8, label="M1 32 workers")
    probe(4_000_000, 256, 16, 8, 4096, label="M1 4M tasks")
