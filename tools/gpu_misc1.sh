#!/bin/bash
OUT=gpurun_out/${1:-misc1}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_operators_gpu.py tests/test_residual_gpu.py tests/test_cli_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -12 | tee $OUT/pytest.log
timeout 200 python tools/e2e_probe.py 8 16 2>&1 | tail -2 | tee $OUT/e2e_probe.txt
timeout 200 python tools/e2e_probe.py 4 16 2>&1 | tail -1 | tee -a $OUT/e2e_probe.txt
timeout 200 python tools/e2e_probe.py 16 32 2>&1 | tail -1 | tee -a $OUT/e2e_probe.txt
ncu --set full --clock-control none --import-source on -k regex:solve_pass3 -s 2 -c 1 -o $OUT/solve3_2048 python tools/profile_stages.py 2048 2048 > $OUT/ncu_full_2048.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:solve_pass3 -s 40 -c 1 -o $OUT/solve3_c2 python tools/profile_step.py c2 0 1 > $OUT/ncu_full_c2.log 2>&1
ls -la $OUT
