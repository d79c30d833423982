#!/bin/bash
# compute-sanitizer passes over a few small ticks (run on the GPU box).  Cooperative launches are replaced by
# plain ones (HQS_DEBUG_NO_COOP=1): the grid is far smaller than the machine, so co-residency still holds.
set -u
export HQS_DEBUG_NO_COOP=1
T="tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_prefill.py"
K="1000-8-6-2 or 4097-16-12-3 or narrow_amounts_with_remainders or eight_resources or indep3_2500_6_5_21 or prefill_steal or min_utilization_moves"
# racecheck / synccheck slow the kernel down by orders of magnitude: the in-kernel time-outs (real clock cycles) then fire,
# which is reported as a failed tick, not as a race
for tool in memcheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest $T -k "$K" -x -q 2>&1 | tail -30
  echo "rc=${PIPESTATUS[0]}"
done
