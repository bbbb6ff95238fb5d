#!/bin/bash
# What the driver runs at round end, in one call: pytest -m gpu (whole suite, one process), smoke, default bench, reference arm.
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>&1; echo "ref rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print('value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', d['e2e'] and round(d['e2e']['value'], 1), 'clocks', d['clocks'])
print('roofline', {k: d['roofline'][k] for k in ('kernel', 'bound', 'achieved', 'peak', 'frac', 'traffic')})
print('cpu_baseline', d['cpu_baseline'])
print({k: v['ms_per_step'] for k, v in d['kernels'].items()})
r = json.loads(open('gpurun_out/bench_reference.json').read().strip().splitlines()[-1])
print('reference', r.get('value'), r.get('unit'), r.get('cpu_baseline', {}).get('cores'))
PY
