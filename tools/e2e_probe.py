"""Breaks the end-to-end M1 step (hqs_ready_push + hqs_tick with pinned host buffers) into its parts."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hyperqueue_b200._lib as L
from workloads import gpu_scheduler
from bench import CFG2, make_workload
from hyperqueue_b200 import priority_from_user

lib = L.load_library()
N, W = 1_000_000, 256
wl = make_workload(CFG2, seed=0)
s = gpu_scheduler(wl, add_tasks=False)
s._sync_classes()
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
handles = pin(np.arange(N, dtype=np.uint32))
cls = pin(wl.task_class.astype(np.uint32))
prio = pin(priority_from_user(wl.task_user_priority))
out = torch.empty(N * 8, dtype=torch.uint8).pin_memory().numpy().view(L.assignment_dtype)
workers = s._worker_structs(0.0)
free = np.ascontiguousarray(wl.worker_free)
total = np.ascontiguousarray(wl.worker_total)
free_after = np.zeros_like(free)
out_n = C.c_uint32(0)


def t(f, n=10):
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


d16 = torch.empty(16 * N, dtype=torch.uint8, device="cuda")
h16 = torch.empty(16 * N, dtype=torch.uint8).pin_memory()
d8 = torch.empty(8 * N, dtype=torch.uint8, device="cuda")
h8 = torch.empty(8 * N, dtype=torch.uint8).pin_memory()
print("raw H2D 16 MB pinned: %.3f ms" % t(lambda: (d16.copy_(h16, non_blocking=True), torch.cuda.synchronize())))
print("raw D2H  8 MB pinned: %.3f ms" % t(lambda: (h8.copy_(d8, non_blocking=True), torch.cuda.synchronize())))
print("host max over 2 x 1M u32: %.3f ms" % t(lambda: (handles.max(), cls.max())))


def push():
    rc = lib.hqs_ready_push_range(s._ctx, 0, N, L.ptr(cls), L.ptr(prio))       # a task array: no handle array over PCIe
    assert rc == 0


def tick():
    rc = lib.hqs_tick(s._ctx, W, L.ptr(workers), L.ptr(free), L.ptr(total), None, N, L.ptr(out), C.byref(out_n),
                      L.ptr(free_after))
    assert rc == 0 and out_n.value == N, (rc, out_n.value)


def both():
    push()
    tick()


both()
tp = []
tt = []
for _ in range(10):
    t0 = time.perf_counter(); push(); t1 = time.perf_counter(); tick(); t2 = time.perf_counter()
    tp.append(t1 - t0); tt.append(t2 - t1)
print("hqs_ready_push_range: %.3f ms   hqs_tick: %.3f ms   sum %.3f ms" % (np.median(tp) * 1e3, np.median(tt) * 1e3,
                                                                     (np.median(tp) + np.median(tt)) * 1e3))
