"""Reads the section cycle sums of the lean first-fit loop from the measuring build (tools/trace_build.sh, -DHQS_TRACE):
   per group: top (prefetch of the next group) and record (group record, loop end); per tile visit: tile load, fit + vote,
   and the whole visit by kind (dead tile / the first worker with room takes all / scan with the tile exhausted / scan
   with the group ending in the tile).
Usage: python tools/trace_probe.py [n_tasks] [n_workers]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from hyperqueue_b200 import _lib as L

L.LIB_PATH = os.path.join(ROOT, "hyperqueue_b200", "libhqsched_b200_trace.so")
import workloads as WL


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    w = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    for scale, tag in [(1024, "M1"), (1, "M2 first tick")]:
        wl = WL.make_independent(n, w, 16, seed=0, free_scale=scale)
        s = WL.gpu_scheduler(wl)
        for it in range(3):
            s.free = wl.worker_free.copy()
            m = s.run_scheduling()
            d = (C.c_uint64 * 8)()
            s._lib.hqs_debug_read(s._ctx, d)
            d = list(d)
            lo = lambda x: x & 0xFFFFFFFF
            kinds = ["dead tile", "first worker takes all", "scan, tile exhausted", "scan, group ends"]
            per = " | ".join(f"{kinds[q]}: {d[2 + q] >> 32} visits {lo(d[2 + q])} cycles" for q in range(4))
            print(f"{tag}: assigned {m.n_assigned()} solver warp {d[6]} cycles | group top {lo(d[0])} | group record {d[0] >> 32} | "
                  f"in visits: tile load {lo(d[1])}, fit + vote {d[1] >> 32} | {per}", flush=True)
            s.rearm()
        s.close()


if __name__ == "__main__":
    main()
