#!/bin/bash
# 2-GPU bench under NCCL settings: how long does the 31 MB gradient all-reduce take?
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus ${NG:-2} --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 0 > gpurun_out/nccl_$tag.json 2> gpurun_out/nccl_$tag.err
  python -c "
import json
d=json.loads(open('gpurun_out/nccl_$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['value'],1), 'img/s', round(d['ms_per_step'],3), 'ms  allreduce', d['kernels']['grad_allreduce']['ms_per_step'])"
}
run default A=1
run ch32 NCCL_MIN_NCHANNELS=32
run ll128 NCCL_PROTO=LL128
run simple NCCL_PROTO=Simple
run nvlsoff NCCL_NVLS_ENABLE=0
run ctas NCCL_MIN_CTAS=32
grep -h "NCCL INFO.*\(channels\|Algo\|NVLS\)" gpurun_out/nccl_default.err | head -5
