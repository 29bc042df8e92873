#!/bin/bash
OUT=gpurun_out/${1:-p3c}; mkdir -p $OUT
timeout 300 python -m pytest tests/test_flow_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -4 | tee $OUT/pytest_quick.log
if grep -q failed $OUT/pytest_quick.log; then
  echo "halo-2 geometry failed; trying FLOW2D_P3_HALO4"
  FLOW2D_P3_HALO4=1 timeout 300 python -m pytest tests/test_flow_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -4 | tee $OUT/pytest_halo4.log
  exit 0
fi
timeout 900 python -m pytest tests/test_stages_gpu.py tests/test_native_sizes_gpu.py tests/test_reference_gpu.py tests/test_ext_gpu.py tests/test_slab_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -4 | tee $OUT/pytest.log
for wl in c2 c4s c4 c3; do
  for v in h2 h4; do
    args="--workload $wl"; [ $wl = c4s ] && args="--workload c4 --streams 1 --pairs 1"
    if [ $v = h4 ]; then export FLOW2D_P3_HALO4=1; else unset FLOW2D_P3_HALO4; fi
    timeout 300 python bench.py $args --steps 5 --warmup 3 --no-extra 2>$OUT/err_${wl}_$v.txt | tail -1 > $OUT/bench_${wl}_$v.json
    python -c "
import json; d=json.load(open('$OUT/bench_${wl}_$v.json')); print('$wl $v value %.1f e2e %.1f ms/step %.3f launch_us %.2f'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['launch_us']))" || tail -3 $OUT/err_${wl}_$v.txt
  done
done
