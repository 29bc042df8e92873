#!/bin/bash
# Short check of the handle-count policy of the cluster solve (the last GPU seconds of round 2).
OUT=gpurun_out/last
mkdir -p $OUT
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.log
timeout 120 python -m pytest tests/test_cluster_gpu.py tests/test_flow_gpu.py tests/test_stages_gpu.py -m gpu -q --timeout 60 2>&1 | tail -6 | tee $OUT/pytest_subset.log
