#!/bin/bash
OUT=gpurun_out/${1:-ncu}
mkdir -p $OUT
ncu --set full --clock-control none --import-source on -k regex:solve_pass -s 40 -c 1 -o $OUT/solve_c2 python tools/profile_step.py c2 0 1 > $OUT/ncu_full_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:solve_pass -s 1 -c 1 -o $OUT/solve_c3first python tools/profile_stages.py 2048 2048 > $OUT/ncu_full_c3.log 2>&1
ls -la $OUT
