#!/bin/bash
# same-box A/B of one environment switch inside the full step: tools/gpu_ab_env.sh KDCC_PW_NO_ASTAT=1
mkdir -p gpurun_out
for rep in 1 2; do for e in A=1 "$1"; do
env $e python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ab_env.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/ab_env.json').read().strip().splitlines()[-1]); k=d['kernels']; print('$e', round(d['value'],1), 'img/s', round(d['ms_per_step'],3), 'ms', d['clocks']['sm_mhz'], 'MHz', ' '.join('%s %.3f'%(n,k[n]['ms_per_step']) for n in ('pw_fwd','pw_bwd_dx','pw_bwd_dw','dw_fwd','dw_bwd','hint_loss')))"
done; done
