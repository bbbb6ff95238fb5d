#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "nchw" -p no:cacheprovider 2>&1 | tail -2
for v in 0; do
KDCC_TC_DEBUG=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ab_$v.json 2>&1
python -c "
import json
d=json.loads(open('gpurun_out/ab_$v.json').read().strip().splitlines()[-1]); print('dbg=$v', round(d['value'],1), 'img/s', {k:v['ms_per_step'] for k,v in d['kernels'].items() if k.startswith('dw')})"
done
