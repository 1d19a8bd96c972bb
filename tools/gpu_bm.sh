#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py tests/test_gpu_reference_loop.py -x -q 2>&1 | tail -3
for wl in bmcnet_nfs plain_nfs; do
timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --only-headline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$wl value', d['value'], 'ms', d['ms_per_step'], 'sustained', d.get('value_sustained'), 'e2e', d['e2e']['value'])"
done
