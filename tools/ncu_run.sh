OUT=gpurun_out/ab2; mkdir -p $OUT
ncu --set full --clock-control none --import-source on -k regex:solve_pass2 -s 40 -c 1 -o $OUT/solve_c2_v2 python tools/profile_step.py c2 0 1 > $OUT/ncu_c2.log 2>&1
python tools/ncu_phases.py $OUT/solve_c2_v2.ncu-rep > $OUT/solve_c2_v2_summary.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:solve_pass2 -s 200 -c 1 -o $OUT/solve_c3_v2 python tools/profile_step.py c3 0 1 > $OUT/ncu_c3.log 2>&1
python tools/ncu_phases.py $OUT/solve_c3_v2.ncu-rep > $OUT/solve_c3_v2_summary.txt 2>&1
head -60 $OUT/solve_c3_v2_summary.txt
