#!/bin/bash
# round 2, first GPU pass: full GPU test suite + the default bench line (all workloads, latency, eager legs)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > gpurun_out/r2_gpu.txt 2>&1
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
        print(k, v['value'], v['e2e']['value'], v.get('value_sustained'))
    print('lat', d.get('latency_b1'))
    print('eager', d.get('eager_b200'))
    print('roof', d['roofline']['frac'], d['roofline_voxel']['frac'], d['roofline_encoder']['frac'])
except Exception as e:
    print('bench parse failed', e)
PY
