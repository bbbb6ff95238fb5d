#!/bin/bash
# What the driver runs at round end, in one call: pytest -m gpu (whole suite, one process), smoke, default bench with the
# driver's flags, reference arm; then the round's ncu evidence (tools/gpu_profile_r2.sh) when PROFILE=1.
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --gpus 1 --steps ${REF_STEPS:-20} --warmup ${REF_WARMUP:-5} > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print('value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', d['e2e'] and round(d['e2e']['value'], 1), 'clocks', d['clocks'])
print('roofline', {k: d['roofline'].get(k) for k in ('kernel', 'bound', 'achieved', 'peak', 'frac', 'traffic', 'smem_floor')})
print('cpu_baseline', d['cpu_baseline'])
print({k: v['ms_per_step'] for k, v in d['kernels'].items()})
print('api', d['api_modules'] and (round(d['api_modules']['value'], 1), d['api_modules']['vs_hotpath']))
print('whole', d['whole_step'] and {k: d['whole_step'].get(k) for k in ('value', 'ms_per_step', 'hot_path_share', 'error')})
print('gpu_baseline', d['gpu_baseline'] and {k: (v['kdcc_ms'], v['speedup_vs_best_stock']) for k, v in d['gpu_baseline'].get('families', {}).items()})
print('k3', d['kernels_k3'])
print('gscnn', d['kernels_gscnn'] and {k: d['kernels_gscnn'][k] for k in ('img_per_s', 'ms_per_step', 'dw_fwd', 'dw_bwd') if k in d['kernels_gscnn']})
r = json.loads(open('gpurun_out/bench_reference.json').read().strip().splitlines()[-1])
print('reference', r.get('value'), r.get('unit'), r.get('cpu_baseline', {}).get('cores'), r.get('whole_step'), 'same config:', r.get('config') == d['config'])
PY
if [ -n "$PROFILE" ]; then bash tools/gpu_profile_r2.sh; fi
