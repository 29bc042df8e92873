#!/bin/bash
OUT=gpurun_out/${1:-misc2}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_flow_gpu.py tests/test_cli_gpu.py tests/test_reference_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -6 | tee $OUT/pytest.log
timeout 200 python tools/e2e_probe.py 8 16 2>&1 | tail -1 | tee $OUT/e2e_probe.txt
timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --no-extra 2>$OUT/err_c4.txt | tail -1 > $OUT/bench_c4.json
python -c "
import json; d=json.load(open('$OUT/bench_c4.json')); print('c4 value %.1f e2e %.1f ms/step %.3f e2e ms %.3f'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step']))"
timeout 300 python tools/bench_sequence.py 32 8 2>&1 | tail -3 | tee $OUT/sequence_h8.txt
