#!/bin/bash
# round-2 check: GPU tests, then the default bench (and optionally the reference arm) -> gpurun_out/
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -${TAILN:-8} | tee gpurun_out/r2_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2_bench.json').read().strip().splitlines()[-1])
    print("value %.1f img/s  %.3f ms/step  e2e %s" % (d['value'], d['ms_per_step'], d['e2e'] and round(d['e2e']['value'], 1)))
    print("clocks", d['clocks'])
    for k, v in d['kernels'].items():
        print("  %-12s %7.3f ms  %s %s frac %s" % (k, v['ms_per_step'], v.get('achieved'), v.get('unit'), v.get('frac')))
    print("api", d.get('api_modules'))
    print("gpu_baseline", json.dumps(d.get('gpu_baseline'))[:3000])
    print("k3", d.get('kernels_k3'))
    print("roofline", d['roofline'])
except Exception as e:
    print("parse failed", e)
PY
