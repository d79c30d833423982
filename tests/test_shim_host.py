"""Host logic of the C++ shim (tako_b200::GpuCore) without a GPU: hyperqueue_b200/csrc/tako_shim.cpp is compiled
against a TEST DOUBLE of the C ABI (tests/mock/fake_hqsched.cpp, a host-memory first-fit) and driven through
interning, batched pushes, cancellation, applying a tick to the worker mirror, resource return, worker removal, and the
bookkeeping of proactive filling: prefill records, retract + redirect, on_retract_response, RunningPrefilled.  The real library is exercised by the same shim on the GPU
(tests/test_gpu_edges.py::test_cpp_shim_selftest)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_host_logic_against_the_abi_double(tmp_path):
    exe = str(tmp_path / "shim_host_test")
    srcs = [os.path.join(ROOT, "tests", "mock", "shim_host_test.cpp"), os.path.join(ROOT, "hyperqueue_b200", "csrc", "tako_shim.cpp"),
            os.path.join(ROOT, "tests", "mock", "fake_hqsched.cpp")]
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-o", exe] + srcs, check=True, cwd=ROOT)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_the_double_is_not_part_of_the_product():
    pkg = os.path.join(ROOT, "hyperqueue_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                assert "fake_hqsched" not in open(os.path.join(dirpath, f), errors="ignore").read(), f
    assert "fake_hqsched" not in open(os.path.join(ROOT, "__graft_entry__.py")).read()
    assert "fake_hqsched" not in open(os.path.join(ROOT, "bench.py")).read()
