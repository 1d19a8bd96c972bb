#!/bin/bash
# quick GPU loop: model parity tests, then short bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
for wl in ${WORKLOADS:-plain_nfs bmcnet_nfs}; do
  timeout 600 python bench.py --workload $wl --steps 50 --warmup 5 --cpu-steps 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"
  tail -c 400 gpurun_out/bench_$wl.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$wl.json'))
    print('$wl value %.0f e2e %.0f ms/step %.3f roofline %.0f TF (%.2f) enc %.0f GB/s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline_encoder']['achieved']))
except Exception as e: print('no json', e)
PY
done
