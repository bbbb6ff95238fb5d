#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --dw k3d1p1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ab_k3.json 2>&1
python -c "
import json
d=json.loads(open('gpurun_out/ab_k3.json').read().strip().splitlines()[-1]); print('k3', round(d['value'],1), 'img/s'); [print('  ',k,v) for k,v in d['kernels'].items()]"
python bench.py --steps 10 --warmup 3 --dw k3d1p1 --layout nhwc --no-cpu-baseline --e2e-steps 0 > gpurun_out/ab_k3_nhwc.json 2>&1
python -c "
import json
d=json.loads(open('gpurun_out/ab_k3_nhwc.json').read().strip().splitlines()[-1]); print('k3 nhwc', round(d['value'],1), 'img/s'); [print('  ',k,v) for k,v in d['kernels'].items() if k.startswith('dw')]"
