"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/hqsched.h declares, struct layouts match the header, and — with no CUDA device — fails loudly
instead of falling back to a CPU path."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from hyperqueue_b200 import _lib
    return _lib.load_library()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "hqsched.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(hqs_[a-z_]+)\s*\(", hdr))
    assert len(declared) >= 18
    from hyperqueue_b200 import _lib
    assert declared == set(_lib.ABI_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.hqs_abi_version() == 1


def test_struct_layouts():
    from hyperqueue_b200 import _lib
    assert C.sizeof(_lib.hqs_variant) == 16 * 8 + 16
    assert C.sizeof(_lib.hqs_class) == 8 + 8 * C.sizeof(_lib.hqs_variant)
    assert C.sizeof(_lib.hqs_worker) == 24 == _lib.worker_dtype.itemsize
    assert _lib.assignment_dtype.itemsize == 8
    assert _lib.assignment_dtype.fields["worker"][1] == 4 and _lib.assignment_dtype.fields["variant"][1] == 6


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    ctx = C.c_void_p()
    rc = lib.hqs_create(C.byref(ctx), 0, 4, 0)
    assert rc == -2 and not ctx.value                       # HQS_E_CUDA
    assert b"no CPU fallback" in lib.hqs_last_error(None)
    from hyperqueue_b200 import GpuScheduler, HqsError
    with pytest.raises(HqsError):
        GpuScheduler(4)


def test_product_does_not_import_oracle():
    """The product path may not import, link or execute anything under oracle/ (or the test-only model)."""
    pkg = os.path.join(ROOT, "hyperqueue_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            src = open(os.path.join(dirpath, f), errors="ignore").read() if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) else ""
            if f.endswith(".py"):
                assert not re.search(r"^\s*(import|from)\s+(oracle|greedy_model|parity)\b", src, flags=re.M), f
            if f.endswith((".cu", ".cuh", ".h", ".cpp", ".hpp")):
                assert not re.search(r"#include\s+[\"<][^\">]*oracle", src), f
    so = os.path.join(pkg, "libhqsched_b200.so")
    out = __import__("subprocess").run(["ldd", so], capture_output=True, text=True).stdout
    assert "hqjudge" not in out and "oracle" not in out


def test_priority_mapping_matches_oracle():
    from hyperqueue_b200 import priority_from_user
    from oracle.model import priority_from_user as ref
    ups = np.array([-2**31, -5, -1, 0, 1, 7, 123, 2**31 - 1])
    got = priority_from_user(ups)
    assert [int(x) for x in got] == [ref(int(u)) for u in ups]
    assert (np.diff(got.astype(np.float64)) > 0).all()


def test_cpp_shim_builds_and_exports_the_reference_interface(lib):
    """The host side above the C ABI is C++ (the reference's is Rust): libhqtako_shim.so links against the C-ABI
    library only and exports tako_b200::GpuCore with the reference's operation names."""
    import subprocess
    from hyperqueue_b200 import _lib
    shim = _lib.load_shim()
    assert hasattr(shim, "hqshim_selftest")
    syms = subprocess.run(["nm", "-DC", _lib.SHIM_PATH], capture_output=True, text=True).stdout
    for name in ("tako_b200::GpuCore::get_or_create_resource_rq_id", "tako_b200::GpuCore::on_new_worker",
                 "tako_b200::GpuCore::add_ready_task", "tako_b200::GpuCore::remove_ready_task",
                 "tako_b200::GpuCore::run_scheduling", "tako_b200::GpuCore::on_task_finished",
                 "tako_b200::GpuCore::block_request"):
        assert name in syms, name
    ldd = subprocess.run(["ldd", _lib.SHIM_PATH], capture_output=True, text=True).stdout
    assert "libhqsched_b200.so" in ldd and "hqjudge" not in ldd
    import torch
    if not torch.cuda.is_available():
        # no device: the shim fails loudly (exception caught by the self-test => one failed check), no CPU path
        assert shim.hqshim_selftest(0, 0) >= 1


def test_solver_shared_memory_budget(lib):
    """Every instance of the tick kernel must leave room for its worst-case MANDATORY dynamic shared memory (1024 workers
    x 16 resource slots x 8 B of free amounts + per-worker words + 4096 group-list entries: ~200 KB) inside the 227 KB
    a CTA may use — otherwise a tick fails on the device, which the CPU-only box would not notice."""
    import subprocess
    from hyperqueue_b200 import _lib
    out = subprocess.run(["cuobjdump", "-res-usage", _lib.LIB_PATH], capture_output=True, text=True).stdout
    statics = [int(m) for blk in re.findall(r"Function [^\n]*tick_k[^\n]*\n[^\n]*", out) for m in re.findall(r"SHARED:(\d+)", blk)]
    assert len(statics) == 6
    # mandatory solver arrays: free amounts [W][RT], per-worker words, per-class words, the group list (12 B per entry).
    # Largest supported corners: 1024 workers x 16 wide resource slots with 4096 list entries, and 8192 entries
    # (HQS_MAX_GROUPS) with <= 8 resource slots; a tick beyond both fails with HQS_E_LIMIT before it is launched
    for W, RT, at, Q, n_pos in [(1024, 16, 8, 4096, 4096), (1024, 8, 8, 4096, 8192), (1024, 16, 4, 4096, 8192)]:
        worst_mandatory = W * RT * at + W * (4 + 8 + 1 + 1 + 2) + Q * 3 + n_pos * 12 + 10 * 16
        assert max(statics) + worst_mandatory <= 227 * 1024, (W, RT, at, n_pos, max(statics), worst_mandatory)
    # the emit step of the worker CTAs: <= 128 KB of counters (+ the group records when they fit) + the segment cache
    assert max(statics) + 128 * 1024 + 8 * 1024 <= 227 * 1024
