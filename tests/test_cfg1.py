"""BASELINE.json configs[0] (10 k independent 1-cpu tasks, 4 workers x 128 cpus): the reference's own CPU-runnable
case.  Oracle and specification drain it tick for tick the same way (512 tasks per tick, 20 ticks); the CUDA path
is held to the same numbers on the GPU."""
import numpy as np
import pytest

import greedy_model as G
import parity as P


def test_cfg1_oracle_and_specification_agree():
    wl = P.make_cfg1()
    o_ticks, o_per = P.oracle_drain(wl)
    m_ticks, m_per = G.model_drain(wl)
    assert o_ticks == m_ticks == 20
    assert o_per == m_per == [512] * 19 + [10_000 - 19 * 512]


@pytest.mark.gpu
def test_cfg1_cuda_drain():
    wl = P.make_cfg1()
    ticks, per = P.gpu_drain(wl)
    assert ticks == 20 and per == [512] * 19 + [10_000 - 19 * 512]
