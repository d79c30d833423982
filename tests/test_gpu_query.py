"""GPU twin of tests/test_query_spec.py: every case of the reference's tests/test_query.rs that the specification answers is
answered again by hqs_query on the device (GpuScheduler.new_worker_query: hypothetical workers, ResourceAmount::MAX partial
descriptors, time limits, min_utilization), and the per-worker counts must equal the specification's."""
import pytest

import test_query_spec as T

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _gpu_backend():
    T._BACKEND = "gpu"
    yield
    T._BACKEND = "spec"


for _name in dir(T):
    if _name.startswith("test_"):
        globals()[_name] = getattr(T, _name)
