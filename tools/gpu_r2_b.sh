#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest_gpu.log 2>&1
tail -5 gpurun_out/r2_pytest_gpu.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
tail -c 600 gpurun_out/r2_bench_default.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2_bench_default.json').read().strip().splitlines()[0])
    print('value', d['value'], 'e2e', d['e2e']['value'], 'sustained', d.get('value_sustained'))
    for k, v in d.get('workloads', {}).items():
        print(k, v['value'], v.get('e2e', {}).get('value'), v.get('value_sustained'), v.get('ms_per_iteration'))
    print('roof', d['roofline']['frac'], d['roofline_voxel']['frac'], d['roofline_encoder']['frac'])
except Exception as e:
    print('bench parse failed', e)
PY
