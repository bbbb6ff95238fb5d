#!/bin/bash
# 2-GPU check of the data-parallel path: peer-push gradient exchange vs the NCCL all-reduce (same box, back to back)
mkdir -p gpurun_out
for mode in ${MODES:-peer nccl}; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus ${NG:-2} --steps 20 --warmup 5 \
    --grad-exchange $mode --no-extras --e2e-steps 0 > gpurun_out/n2_$mode.json 2> gpurun_out/n2_$mode.err; echo "$mode rc=$?"
  python -c "
import json; d=json.loads(open('gpurun_out/n2_$mode.json').read().strip().splitlines()[-1]); print('$mode', round(d['value'],1), round(d['ms_per_step'],3)); print(d['grad_exchange']); print(d['scaling_diag']); print(d['kernels'].get('grad_allreduce'), d['kernels'].get('optimizer')); print(d['losses'])"
  grep -v Warning gpurun_out/n2_$mode.err | tail -4
done
