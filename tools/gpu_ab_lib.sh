#!/bin/bash
# same-box A/B of two builds of the library inside the full step: tools/gpu_ab_lib.sh <variant tag>
mkdir -p gpurun_out
P=$PWD/knowledge-distillation-by-replacing-cheap-conv_b200
for rep in 1 2; do for lib in libkdcc.so libkdcc_$1.so; do
KDCC_LIB=$P/$lib python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-extras --e2e-steps 0 > gpurun_out/ab_$lib.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/ab_$lib.json').read().strip().splitlines()[-1]); k=d['kernels']; print('$lib', round(d['value'],1), 'img/s', round(d['ms_per_step'],3), 'ms', d['clocks']['sm_mhz'], 'MHz', ' '.join('%s %.3f'%(n,k[n]['ms_per_step']) for n in ('pw_fwd','pw_bwd_dx','pw_bwd_dw','dw_fwd','dw_bwd')))"
done; done
