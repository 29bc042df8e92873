#!/bin/bash
OUT=gpurun_out/${1:-misc3}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_flow_gpu.py tests/test_ext_gpu.py tests/test_cli_gpu.py tests/test_residual_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -4 | tee $OUT/pytest.log
timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --no-extra 2>$OUT/err_c4.txt | tail -1 > $OUT/bench_c4.json
python -c "
import json; d=json.load(open('$OUT/bench_c4.json')); print('c4 value %.1f e2e %.1f ms/step %.3f e2e ms %.3f'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step']))"
timeout 300 python tools/bench_sequence.py 128 8 2>/dev/null | tail -1 > $OUT/sequence_h8.json
python -c "
import json; d=json.load(open('$OUT/sequence_h8.json')); print({k:(v['pairs_per_s'], v['mpix_per_s']) for k,v in d['inputs'].items()})"
