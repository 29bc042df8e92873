#!/bin/bash
OUT=gpurun_out/r01f
mkdir -p $OUT
timeout 300 python -m pytest tests/test_residual_gpu.py -m gpu -q -x 2>&1 | tail -15 > $OUT/pytest_res.log; tail -15 $OUT/pytest_res.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > $OUT/pytest.log; tail -3 $OUT/pytest.log
