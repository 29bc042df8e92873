#!/bin/bash
# Thread-block-cluster solve: parity tests first, then the A/B of every switch on the same box.
#   gpurun --timeout 420 -- bash tools/gpu_cluster.sh
OUT=gpurun_out/cluster
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > $OUT/gpu.txt
timeout 260 python -m pytest tests/test_cluster_gpu.py -q --timeout 90 -x 2>&1 | tail -25 > $OUT/pytest_cluster.log
tail -4 $OUT/pytest_cluster.log
if grep -qE "failed|error|Timeout" $OUT/pytest_cluster.log || ! grep -q " passed" $OUT/pytest_cluster.log; then
  echo "cluster parity tests did not pass: no A/B"
  # which part is broken?  whole-level against the oracle, one case at a time
  timeout 100 python -m pytest tests/test_cluster_gpu.py -q --timeout 30 -k "whole_level or shapes or pass_equals" 2>&1 | tail -40 > $OUT/pytest_cluster_all.log
  tail -5 $OUT/pytest_cluster_all.log
  exit 0
fi
timeout 240 python tools/ab_cluster.py $OUT/ab.json 2> $OUT/ab.log > /dev/null
tail -45 $OUT/ab.log
