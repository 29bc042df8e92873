#!/bin/bash
# slab tests only, with tight time limits (a wiring mistake must not eat the GPU budget)
OUT=gpurun_out/${1:-slab}; mkdir -p $OUT
timeout 400 python -m pytest tests/test_slab_gpu.py tests/test_slab_multigpu.py -m gpu -q -x 2>&1 | tail -30 > $OUT/pytest_slab.log; tail -25 $OUT/pytest_slab.log
