#!/bin/bash
OUT=gpurun_out/r01j
mkdir -p $OUT
timeout 1500 python tools/stress_parity.py 1000 2 > $OUT/stress_1000.log 2>&1; tail -15 $OUT/stress_1000.log
