"""Section cycle sums of the WIDE first-fit loop (warp 0) from the measuring build (tools/trace_build.sh, -DHQS_TRACE).
Usage: python tools/trace_wide.py [n_tasks] [n_workers]      (HQS_LIB selects another trace build)"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from hyperqueue_b200 import _lib as L

L.LIB_PATH = os.path.join(ROOT, "hyperqueue_b200", os.environ.get("HQS_LIB", "libhqsched_b200_trace.so"))
import workloads as WL

NAMES = ["top", "fit", "sum+vote+record", "exchange", "warp prefix", "scan", "takes+segments", "group record"]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    w = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    wl = WL.make_independent(n, w, 16, seed=0, free_scale=1024)
    s = WL.gpu_scheduler(wl)
    for it in range(3):
        s.free = wl.worker_free.copy()
        m = s.run_scheduling()
        d = (C.c_uint64 * 8)()
        s._lib.hqs_debug_read(s._ctx, d)
        d = list(d)
        sec = []
        for q in range(4):
            sec += [d[q] & 0xFFFFFFFF, d[q] >> 32]
        steps = max(d[4], 1)
        print(f"M1 n={n} w={w}: assigned {m.n_assigned()} solver warp {d[6]} cycles, {d[4]} steps | " +
              " | ".join(f"{NAMES[q]} {sec[q]} ({sec[q] // steps}/step)" for q in range(8)), flush=True)
        s.rearm()
    s.close()


if __name__ == "__main__":
    main()
