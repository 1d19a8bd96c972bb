#!/bin/bash
# whole GPU suite + the driver's default bench command on one GPU
mkdir -p gpurun_out
S=$(date +%s)
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$? wall $(( $(date +%s) - S )) s"
tail -4 gpurun_out/r2_pytest_gpu.log
S=$(date +%s)
timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench rc=$? wall $(( $(date +%s) - S )) s"
tail -c 300 gpurun_out/r2_bench_default.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_default.json') if l.startswith('{')][-1])
print('value %.0f sustained %.0f e2e %.0f ms %.3f roofline %.3f enc %.2f vox %.2f launches %s' % (d['value'], d.get('value_sustained', 0), d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline_encoder']['frac'], d['roofline_voxel']['frac'], d['gpu_launches']))
for k, v in d['workloads'].items():
    print(k, {kk: (round(vv, 1) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ('value', 'value_sustained', 'ms_per_step', 'ms_per_iteration', 'ms_per_iteration_eager', 'batch_per_gpu', 'loss_finite', 'batch_8')})
print('cpu', d['cpu_baseline'], 'eager', d.get('eager_b200'), 'lat', d.get('latency_b1'))
PY
